// kfb_icp.cu -- projective-data-association point-to-plane ICP (replaces kf::device::rigidICP,
// kfusion/src/rigid_icp.cu:46-169, and the device side of ICPRegistration::rigidTransform's loop,
// kfusion/src/icp_registration.cpp:21-43).
//
// One accumulation does correspondence search, gating, the residual row [s x n, n, n.(d-s)] and the 27 unique
// products of the 6x7 normal equations, reduces them with a transposing warp exchange, one shared-memory stage
// per block and a single-pass grid reduction (last-block-done ticket, fixed summation order => bit-reproducible),
// and writes the 27 sums as tagged 16-byte chunks straight into mapped pinned host memory.  The host spins on
// the tags: no cudaMemcpy, no allocation, no second kernel, no device sync per iteration (the reference does
// 2 launches + 2 cudaMalloc + 2 cudaFree + a blocking memcpy).  Two drivers share the code:
// icp_kernel (one launch per kfb_icp_accumulate) and icp_freerun_kernel (the whole schedule in one launch, the
// host checking every pose the kernel computed for itself -- see the comment at that kernel).
//
// Numerics: products are f32 exactly as in the reference (`smem[tid] = row[i]*row[j]`),
// sums are carried in f64 end to end (the reference rounds per-32x32-tile sums to f32 in between; the
// difference is ~1e-7 relative per entry and far below the 1e-4 pose tolerance).  Coverage: compat_icp_rows
// reproduces the reference's truncated grid floor(w/32) x floor(h/32) tiles (§9 Q7).
#include "kfb_common.cuh"
#include <xmmintrin.h>
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace kfb
{

struct IcpArgs
{
    const float4 *cur_v, *cur_n, *pre_v, *pre_n;
    Intr k;
    int cov_w, cov_h; // pixels visited: [0,cov_w) x [0,cov_h)
    int stride;       // threads of the full grid (SMs x ICP_THREADS): the pixel -> thread map does not depend on how many CTAs take part
    Pose pose;
    float dist_thres, sine_thres;
    double *partials;           // [gridDim.x][27]
    unsigned int *ticket;
    IcpHostResult *out;         // mapped host memory
    unsigned long long seq;
};

#define ICP_THREADS 480                 // threads of a CTA (15 warps)
#define ICP_BAR() asm volatile("bar.sync 1, 480;" ::: "memory") // barrier over the workers only

// findCoresp (rigid_icp.cu:46-80) + row (rigid_icp.cu:85-95), split so that the loads of several pixels can
// be in flight together: (1) current vertex/normal -> transformed point s and the model pixel it projects to,
// (2) the two model-map gathers, (3) gates + row.
struct IcpProbe
{
    float sx, sy, sz;
    float4 nc4;
    int j; // model pixel, -1 = no correspondence
};
__device__ __forceinline__ void icp_probe_values(const IcpArgs &a, const float4 nc4, const float4 vc4, IcpProbe &o);
__device__ __forceinline__ void icp_probe(const IcpArgs &a, int p, int npix, IcpProbe &o)
{
    o.j = -1;
    if (p >= npix) return;
    const int y = p / a.cov_w, x = p - y * a.cov_w;
    const int i = y * a.k.w + x;
    const float4 nc4 = __ldg(a.cur_n + i);
    icp_probe_values(a, nc4, __ldg(a.cur_v + i), o);
}
// the same from values the caller already holds (o.j must be -1 on entry)
__device__ __forceinline__ void icp_probe_values(const IcpArgs &a, const float4 nc4, const float4 vc4, IcpProbe &o)
{
    o.nc4 = nc4;
    if (isnan(o.nc4.x)) return;
    const float3 r = rot3(a.pose.R, vc4.x, vc4.y, vc4.z);
    o.sx = __fadd_rn(r.x, a.pose.t[0]); o.sy = __fadd_rn(r.y, a.pose.t[1]); o.sz = __fadd_rn(r.z, a.pose.t[2]);
    // Intrs::proj (device_utils.cuh:15-21)
    const float qx = __fdividef(o.sx, o.sz), qy = __fdividef(o.sy, o.sz);
    const int px = __float2int_rn(__fmaf_rn(qx, a.k.fx, a.k.cx));
    const int py = __float2int_rn(__fmaf_rn(qy, a.k.fy, a.k.cy));
    if (!(o.sz > 0.f && px >= 0 && py >= 0 && px < a.k.w && py < a.k.h)) return;
    o.j = py * a.k.w + px;
}
__device__ __forceinline__ bool icp_row(const IcpArgs &a, const IcpProbe &o, const float4 vp, const float4 np, float row[7])
{
    const float sx = o.sx, sy = o.sy, sz = o.sz;
    const float dx = __fsub_rn(sx, vp.x), dy = __fsub_rn(sy, vp.y), dz = __fsub_rn(sz, vp.z);
    const float dist = __fsqrt_rn(dot3c(dx, dy, dz, dx, dy, dz));
    if (!(dist <= a.dist_thres)) return false;
    const float3 nc = rot3(a.pose.R, o.nc4.x, o.nc4.y, o.nc4.z);
    const float cx = __fmaf_rn(nc.y, np.z, -__fmul_rn(nc.z, np.y));
    const float cy = __fmaf_rn(nc.z, np.x, -__fmul_rn(nc.x, np.z));
    const float cz = __fmaf_rn(nc.x, np.y, -__fmul_rn(nc.y, np.x));
    const float sine = __fsqrt_rn(dot3c(cx, cy, cz, cx, cy, cz));
    if (!(sine <= a.sine_thres)) return false;
    row[0] = __fmaf_rn(sy, np.z, -__fmul_rn(sz, np.y));
    row[1] = __fmaf_rn(sz, np.x, -__fmul_rn(sx, np.z));
    row[2] = __fmaf_rn(sx, np.y, -__fmul_rn(sy, np.x));
    row[3] = np.x; row[4] = np.y; row[5] = np.z;
    const float ex = __fsub_rn(vp.x, sx), ey = __fsub_rn(vp.y, sy), ez = __fsub_rn(vp.z, sz);
    row[6] = __fmaf_rn(np.z, ez, __fmaf_rn(np.x, ex, __fmul_rn(np.y, ey)));
    return true;
}

__device__ __forceinline__ float4 ld_volatile_f4(const volatile float *p)
{
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_f4(float *p, const float4 v)
{
    asm volatile("st.volatile.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// ---- shared pieces ------------------------------------------------------------------------------------
struct IcpLevel
{
    const float4 *cur_v, *cur_n, *pre_v, *pre_n;
    Intr k;
    int cov_w, cov_h;
    int nact; // CTAs that own pixels at this level (see icp_setup); thread t of CTA c visits pixels c * 480 + t + k * nact * 480
};

#define ICP_BATCH 4
// A thread visits the same pixels in every iteration of a level (fixed pixel -> thread map), and the current
// frame's vertex / normal of a pixel do not change while the pose does: the whole-schedule kernel keeps the first
// ICP_CACHE_SLOTS pixels of every thread in shared memory (thread-private slots, no synchronisation), which takes
// one L2 round trip out of each batch's dependent chain (current maps -> projection -> model gathers).
#define ICP_CACHE_SLOTS 5 // 640x480 on 148 SMs: 4.3 pixels per thread; 2 x 16 B x 5 x 480 = 75 KB per CTA
template <bool CACHED>
__device__ __forceinline__ void icp_accumulate_pixels(const IcpArgs &a, double acc[27], int first, int stride, float4 *cache = nullptr,
                                                      bool fill = false)
{
    const int npix = a.cov_w * a.cov_h;
    int slot0 = 0;
    for (int p0 = first; p0 < npix; p0 += ICP_BATCH * stride, slot0 += ICP_BATCH)
    {
        IcpProbe pr[ICP_BATCH];
        float4 vp[ICP_BATCH], np[ICP_BATCH];
#pragma unroll
        for (int b = 0; b < ICP_BATCH; ++b)
        {
            if (!CACHED) icp_probe(a, p0 + b * stride, npix, pr[b]);
            else
            {
                const int p = p0 + b * stride, slot = slot0 + b;
                pr[b].j = -1;
                if (p < npix)
                {
                    float4 *c = cache + (size_t)(2 * slot) * ICP_THREADS + threadIdx.x;
                    if (slot < ICP_CACHE_SLOTS && !fill) icp_probe_values(a, c[0], c[ICP_THREADS], pr[b]);
                    else
                    {
                        const int y = p / a.cov_w, x = p - y * a.cov_w;
                        const int i = y * a.k.w + x;
                        const float4 nc4 = __ldg(a.cur_n + i), vc4 = __ldg(a.cur_v + i);
                        if (slot < ICP_CACHE_SLOTS) { c[0] = nc4; c[ICP_THREADS] = vc4; }
                        icp_probe_values(a, nc4, vc4, pr[b]);
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < ICP_BATCH; ++b)
        {
            const int j = max(pr[b].j, 0);
            vp[b] = __ldg(a.pre_v + j);
            np[b] = __ldg(a.pre_n + j);
        }
        // pixels are accumulated in increasing p within a thread: the order is fixed => reproducible sums
#pragma unroll
        for (int b = 0; b < ICP_BATCH; ++b)
        {
            float row[7];
            if (pr[b].j >= 0 && icp_row(a, pr[b], vp[b], np[b], row))
            {
                int s = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = i; j < 7; ++j) acc[s++] += (double)__fmul_rn(row[i], row[j]);
            }
        }
    }
}
// warp reduction by transposition (lane l ends up owning value l: 31 exchanges instead of 27 five-step trees),
// then one shared stage; threads 0..26 end up with the block's sums (valid for threadIdx.x < 27)
__device__ __forceinline__ double icp_block_reduce(double acc[27], double (*sm)[27])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = i < 27 ? acc[i] : 0.0;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1)
    {
        const bool up = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i)
        {
            const double send = up ? v[i] : v[i + half];
            const double keep = up ? v[i + half] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
        }
    }
    if (lane < 27) sm[warp][lane] = v[0];
    ICP_BAR();
    double s = 0.0;
    if (threadIdx.x < 27)
    {
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; ++w) s += sm[w][threadIdx.x];
    }
    return s;
}
// last block: fixed-order sum of the per-block partials.  Thread t owns value v = t & 31 over the block
// slice {t >> 5, t >> 5 + 8, ...}: all loads of a thread are independent (one L2 round trip), the summation
// order is a fixed function of the grid size => bit-reproducible.  Lanes 0..26 of warp 0 return the totals.
__device__ __forceinline__ double icp_final_reduce(const double *partials, int nb, double (*red)[28])
{
    const int v = threadIdx.x & 31, slice = threadIdx.x >> 5;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (v < 27)
    {
        const int stride = ICP_THREADS / 32;
        int bb = slice;
        for (; bb + 3 * stride < nb; bb += 4 * stride)
        {
            const double p0 = __ldcg(partials + (size_t)bb * 27 + v);
            const double p1 = __ldcg(partials + (size_t)(bb + stride) * 27 + v);
            const double p2 = __ldcg(partials + (size_t)(bb + 2 * stride) * 27 + v);
            const double p3 = __ldcg(partials + (size_t)(bb + 3 * stride) * 27 + v);
            s0 += p0; s1 += p1; s2 += p2; s3 += p3;
        }
        for (; bb < nb; bb += stride) s0 += __ldcg(partials + (size_t)bb * 27 + v);
        red[slice][v] = (s0 + s1) + (s2 + s3);
    }
    ICP_BAR();
    double fin = 0.0;
    if (threadIdx.x < 27)
    {
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; ++w) fin += red[w][threadIdx.x];
    }
    return fin;
}
// result chunk i = {sum_i, tag}: one aligned 16-byte store per lane, the tag (sequence number) travels with
// the value, so the host needs no separate flag and the device no system fence
__device__ __forceinline__ void icp_post(IcpHostResult::Chunk *chunks, int i, double v, unsigned long long seq)
{
    asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(chunks + i), "l"(__double_as_longlong(v)), "l"(seq) : "memory");
}

// one accumulation, pose by parameter (kfb_icp_accumulate)
__global__ void __launch_bounds__(ICP_THREADS) icp_kernel(const IcpArgs a)
{
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.0;
    icp_accumulate_pixels<false>(a, acc, blockIdx.x * ICP_THREADS + threadIdx.x, a.stride);
    __shared__ double sm[ICP_THREADS / 32][27];
    __shared__ double red[ICP_THREADS / 32][28];
    __shared__ bool is_last;
    const double s = icp_block_reduce(acc, sm);
    if (threadIdx.x < 27)
    {
        a.partials[(size_t)blockIdx.x * 27 + threadIdx.x] = s;
        __threadfence();
    }
    ICP_BAR();
    if (threadIdx.x == 0)
    {
        const unsigned int t = atomicInc(a.ticket, gridDim.x - 1); // wraps to 0 on the last block
        is_last = (t == gridDim.x - 1);
    }
    ICP_BAR();
    if (!is_last) return;
    __threadfence();
    const double fin = icp_final_reduce(a.partials, (int)gridDim.x, red);
    if (threadIdx.x < 27) icp_post(a.out->chunk, threadIdx.x, fin, a.seq);
}

// the host's ICPRegistration::solve (LDL^T branch) + Tinc + camera_pose * Tinc (kfusion/src/icp_registration.cpp,
// cvlite.hpp), operation for operation.  false: the system is not numerically positive definite (no prediction).
__device__ __noinline__ bool icp_predict_pose(const double *in27, const float *cur, float *out)
{
    double A[6][6], b[6], L[6][6], d[6], inv[6], t[6][6], z[6], x[6];
    {
        int s = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 7; ++j)
            {
                const double v = in27[s++];
                if (j == 6) b[i] = v;
                else A[i][j] = A[j][i] = v;
            }
    }
    // A = L D L^T with one reciprocal per column, the host's sequence (icp_registration.cpp: solve)
#pragma unroll
    for (int j = 0; j < 6; ++j)
    {
        double dj = A[j][j];
#pragma unroll
        for (int q = 0; q < j; ++q) { t[j][q] = __dmul_rn(L[j][q], d[q]); dj = __dsub_rn(dj, __dmul_rn(L[j][q], t[j][q])); }
        if (!(dj > 0.0)) return false;
        d[j] = dj;
        inv[j] = __drcp_rn(dj);
#pragma unroll
        for (int i = j + 1; i < 6; ++i)
        {
            double sum = A[i][j];
#pragma unroll
            for (int q = 0; q < j; ++q) sum = __dsub_rn(sum, __dmul_rn(L[i][q], t[j][q]));
            L[i][j] = __dmul_rn(sum, inv[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
    {
        double sum = b[i];
#pragma unroll
        for (int q = 0; q < i; ++q) sum = __dsub_rn(sum, __dmul_rn(L[i][q], z[q]));
        z[i] = sum;
    }
#pragma unroll
    for (int i = 5; i >= 0; --i)
    {
        double sum = __dmul_rn(z[i], inv[i]);
#pragma unroll
        for (int q = i + 1; q < 6; ++q) sum = __dsub_rn(sum, __dmul_rn(L[q][i], x[q]));
        x[i] = sum;
    }
    // Tinc = Affine3f(rvec = (float)x[0..2], t = (float)x[3..5]): Rodrigues in double, stored as float
    float T[12] = {1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f};
    const double rx = (double)(float)x[0], ry = (double)(float)x[1], rz = (double)(float)x[2];
    const double theta = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)));
    if (theta >= 2.220446049250313e-16)
    {
        double sn, c;
        sincos(theta, &sn, &c);
        const double c1 = __dsub_rn(1.0, c), it = __drcp_rn(theta);
        const double ux = __dmul_rn(rx, it), uy = __dmul_rn(ry, it), uz = __dmul_rn(rz, it);
        const double xx = __dmul_rn(__dmul_rn(c1, ux), ux), yy = __dmul_rn(__dmul_rn(c1, uy), uy), zz = __dmul_rn(__dmul_rn(c1, uz), uz);
        const double xy = __dmul_rn(__dmul_rn(c1, ux), uy), xz = __dmul_rn(__dmul_rn(c1, ux), uz), yz = __dmul_rn(__dmul_rn(c1, uy), uz);
        const double sx = __dmul_rn(sn, ux), sy = __dmul_rn(sn, uy), sz = __dmul_rn(sn, uz);
        T[0] = (float)__dadd_rn(c, xx);  T[1] = (float)__dsub_rn(xy, sz); T[2] = (float)__dadd_rn(xz, sy);
        T[4] = (float)__dadd_rn(xy, sz); T[5] = (float)__dadd_rn(c, yy);  T[6] = (float)__dsub_rn(yz, sx);
        T[8] = (float)__dsub_rn(xz, sy); T[9] = (float)__dadd_rn(yz, sx); T[10] = (float)__dadd_rn(c, zz);
    }
    T[3] = (float)x[3]; T[7] = (float)x[4]; T[11] = (float)x[5];
    // camera_pose * Tinc (cvlite.hpp operator*): s = 0; s += a(i,q) * b(q,j) for q = 0..2; j == 3: s += a(i,3)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j)
        {
            float acc = 0.f;
            for (int q = 0; q < 3; ++q) acc = __fadd_rn(acc, __fmul_rn(cur[4 * i + q], T[4 * q + j]));
            if (j == 3) acc = __fadd_rn(acc, cur[4 * i + 3]);
            out[4 * i + j] = acc;
        }
    return true;
}

// ---- the whole coarse-to-fine loop in ONE free-running kernel ----------------------------------------------
// kfb_icp_begin/step/end: the grid (one CTA per SM, all co-resident, cooperative launch) runs every iteration of
// the schedule and nobody waits for anybody's permission:
//   * every CTA publishes its 27 partial sums as self-validating 16-byte chunks {value, sequence number ^ value bits} (one
//     trip to L2), then EVERY CTA reads all partials back (one more), sums them in the ordinary kernel's fixed
//     order and computes the next pose with the host's operations in the host's order (icp_predict_pose: IEEE
//     double add/mul/div/sqrt are exactly rounded on both sides; only sin/cos may differ in the last bit, which
//     survives the cast to float with probability ~2^-29) -- identical inputs, identical code, identical pose in
//     every CTA, so the grid needs no ticket, no fence, no release;
//   * CTA 0 posts each iteration's sums TOGETHER WITH THE POSE IT USED into that iteration's slot in mapped host
//     memory and goes on.  The host (icp_step) takes the slot, compares the pose with its own, bit for bit, and
//     only then uses the sums; a pose that differs (a caller with another update rule, sin/cos one ulp apart) ends
//     the free run for that schedule: the remaining iterations are ordinary launches with the caller's poses.
//     The host's solve stays authoritative; the host never writes to the device, the device never reads the host.
// Round 1's kernel had one CTA reduce, predict, wait for the host's verdict on the previous pose and release the
// others (store + fence + ticket + loads + release store + poll: five trips through L2 and a PCIe read per
// iteration): 198 us per frame against 140 us for this one, same box, same poses (profiles/r02_experiments.md).
// Partials are double-buffered by iteration parity: a CTA can only be one iteration ahead of the slowest one
// (it needs that CTA's partial of iteration k + 1 before it can post k + 2).  CTAs without pixels at a level
// sleep until CTA 0 publishes the first pose of the next level they own pixels at (IcpDevGate, tagged chunks).
// Every poll is bounded (KFB_ICP_GATE_TIMEOUT_NS, KFB_ICP_TIMEOUT_NS=<ns> overrides): a CTA that gives up leaves,
// the others follow, the host finds the stream idle without its result and finishes with ordinary launches.
struct IcpFreeArgs
{
    IcpLevel lv[KFB_MAX_LEVELS];
    int iters[KFB_MAX_LEVELS];
    int levels, total;
    float dist_thres, sine_thres;
    IcpTagged *tagged;   // [2][cap][27]
    int cap;
    IcpHostSlot *slots;  // mapped host memory, slot k = iteration k
    IcpDevGate *devgate;
    IcpHostResult *dbg;  // debug stamps
    unsigned long long seq0;
    unsigned long long timeout_ns;
    float pose0[12];
};
__device__ __forceinline__ int icp_level_of(const int *iters, int levels, int k)
{
    int level = levels - 1;
    while (level > 0 && k >= iters[level]) { k -= iters[level]; --level; }
    return level;
}
__device__ __forceinline__ void ld_volatile_tagged(const IcpTagged *p, unsigned long long &v, unsigned long long &t)
{
    asm volatile("ld.volatile.global.v2.b64 {%0, %1}, [%2];" : "=l"(v), "=l"(t) : "l"(p) : "memory");
}
// a device-gate chunk {x, y, z, tag ^ bits(x) ^ bits(y) ^ bits(z)}: the tag validates the payload it travels with
__device__ __forceinline__ float4 gate_chunk(float x, float y, float z, unsigned int tag)
{
    return make_float4(x, y, z, __uint_as_float(tag ^ __float_as_uint(x) ^ __float_as_uint(y) ^ __float_as_uint(z)));
}
__device__ __forceinline__ unsigned int gate_tag(const float4 c)
{
    return __float_as_uint(c.w) ^ __float_as_uint(c.x) ^ __float_as_uint(c.y) ^ __float_as_uint(c.z);
}
#define ICP_FREE_MAXM 10 // partials per thread of the final sum: 15 slices x 10 = 150 CTAs
// icp_final_reduce over tagged partials: the same association (four running sums over the slice, then the slices in
// order).  All of a thread's chunks are loaded unconditionally and together (one trip to L2 per attempt; slots past
// the thread's share repeat its last chunk), and reloaded until every tag is the iteration's.
template <int MAXM>
__device__ __forceinline__ void icp_slice_sum_tagged(const IcpTagged *part, int nb, unsigned long long seq, double (*red)[28],
                                                     unsigned long long timeout_ns, volatile int *fail)
{
    const int v = threadIdx.x & 31, slice = threadIdx.x >> 5;
    const int stride = ICP_THREADS / 32;
    if (v < 27 && slice < nb)
    {
        const int M = (nb - slice + stride - 1) / stride;
        const IcpTagged *mine = part + (size_t)slice * 27 + v;
        unsigned long long p[MAXM], tt[MAXM];
        const unsigned long long t0 = globaltimer_ns();
        bool ok;
        for (;;)
        {
#pragma unroll
            for (int m = 0; m < MAXM; ++m) ld_volatile_tagged(mine + (size_t)(min(m, M - 1) * stride) * 27, p[m], tt[m]);
            ok = true;
#pragma unroll
            for (int m = 0; m < MAXM; ++m) ok = ok && ((tt[m] ^ p[m]) == seq); // tag = seq ^ value bits: a torn chunk cannot pass
            if (ok) break;
            if (globaltimer_ns() - t0 > timeout_ns || *fail) { *fail = 1; break; }
        }
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        const int G4 = M & ~3;
#pragma unroll
        for (int m = 0; m < MAXM; ++m)
        {
            const double x = (ok && m < M) ? __longlong_as_double((long long)p[m]) : 0.0;
            if (m < M)
            {
                if (m >= G4 || (m & 3) == 0) s0 += x;
                else if ((m & 3) == 1) s1 += x;
                else if ((m & 3) == 2) s2 += x;
                else s3 += x;
            }
        }
        red[slice][v] = (s0 + s1) + (s2 + s3);
    }
    else if (v < 27) red[slice][v] = 0.0;
}
__device__ __forceinline__ double icp_final_reduce_tagged(const IcpTagged *part, int nb, unsigned long long seq, double (*red)[28],
                                                         unsigned long long timeout_ns, volatile int *fail)
{
    if (nb <= 3 * (ICP_THREADS / 32)) icp_slice_sum_tagged<3>(part, nb, seq, red, timeout_ns, fail);
    else icp_slice_sum_tagged<ICP_FREE_MAXM>(part, nb, seq, red, timeout_ns, fail);
    ICP_BAR();
    double fin = 0.0;
    if (threadIdx.x < 27)
    {
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; ++w) fin += red[w][threadIdx.x];
    }
    return fin;
}

__global__ void __launch_bounds__(ICP_THREADS) icp_freerun_kernel(const IcpFreeArgs P)
{
    __shared__ double sm[ICP_THREADS / 32][27];
    __shared__ double red[ICP_THREADS / 32][28];
    __shared__ double fin27[27];
    __shared__ int s_fail, s_pred;
    __shared__ float pose_buf[2][12];
    extern __shared__ float4 cur_cache[]; // [ICP_CACHE_SLOTS][normal, vertex][ICP_THREADS]

    IcpArgs a;
    a.dist_thres = P.dist_thres; a.sine_thres = P.sine_thres;
    if (threadIdx.x < 12) pose_buf[0][threadIdx.x] = P.pose0[threadIdx.x];
    if (threadIdx.x == 0) s_fail = 0;
    int k = 0, cached_level = -1;
    ICP_BAR();
    for (;;)
    {
        const int level = icp_level_of(P.iters, P.levels, k);
        const IcpLevel &L = P.lv[level];
        const int nact = L.nact;
        if ((int)blockIdx.x >= nact)
        {
            // no pixels of this level are ours: sleep until the first iteration of a level where some are
            int k2 = k + 1;
            while (k2 < P.total && (int)blockIdx.x >= P.lv[icp_level_of(P.iters, P.levels, k2)].nact) ++k2;
            if (k2 >= P.total) return;
            if (threadIdx.x == 0)
            {
                const unsigned int want = (unsigned int)(P.seq0 + (unsigned long long)k2 + 1ull);
                const unsigned long long t0 = globaltimer_ns();
                float4 c0, c1, c2, c3;
                for (;;)
                {
                    c3 = ld_volatile_f4(P.devgate->chunk + 12);
                    if (gate_tag(c3) == want)
                    {
                        c0 = ld_volatile_f4(P.devgate->chunk);
                        c1 = ld_volatile_f4(P.devgate->chunk + 4);
                        c2 = ld_volatile_f4(P.devgate->chunk + 8);
                        if (gate_tag(c0) == want && gate_tag(c1) == want && gate_tag(c2) == want) break;
                    }
                    else __nanosleep(400);
                    if (globaltimer_ns() - t0 > 3ull * P.timeout_ns) { s_fail = 1; break; }
                }
                if (!s_fail)
                {
                    float *sp = pose_buf[k2 & 1];
                    sp[0] = c0.x; sp[1] = c0.y; sp[2] = c0.z; sp[3] = c3.x;
                    sp[4] = c1.x; sp[5] = c1.y; sp[6] = c1.z; sp[7] = c3.y;
                    sp[8] = c2.x; sp[9] = c2.y; sp[10] = c2.z; sp[11] = c3.z;
                }
            }
            ICP_BAR();
            if (s_fail) return;
            k = k2;
            continue;
        }
        const float *spose = pose_buf[k & 1];
        float *npose = pose_buf[(k + 1) & 1];
        a.cur_v = L.cur_v; a.cur_n = L.cur_n; a.pre_v = L.pre_v; a.pre_n = L.pre_n;
        a.k = L.k; a.cov_w = L.cov_w; a.cov_h = L.cov_h;
        const unsigned long long seq = P.seq0 + (unsigned long long)k + 1ull;
#pragma unroll
        for (int i = 0; i < 9; ++i) a.pose.R.m[i] = spose[4 * (i / 3) + (i % 3)];
#pragma unroll
        for (int i = 0; i < 3; ++i) a.pose.t[i] = spose[4 * i + 3];
        const unsigned long long ts0 = globaltimer_ns();
        double acc[27];
#pragma unroll
        for (int i = 0; i < 27; ++i) acc[i] = 0.0;
        icp_accumulate_pixels<true>(a, acc, blockIdx.x * ICP_THREADS + threadIdx.x, nact * ICP_THREADS, cur_cache, level != cached_level);
        cached_level = level;
        const unsigned long long ts1 = globaltimer_ns();
        const double s = icp_block_reduce(acc, sm);
        IcpTagged *part = P.tagged + (size_t)(k & 1) * 27 * (size_t)P.cap;
        if (threadIdx.x < 27)
            asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(part + (size_t)blockIdx.x * 27 + threadIdx.x),
                         "l"(__double_as_longlong(s)), "l"(seq ^ (unsigned long long)__double_as_longlong(s)) : "memory");
        const double fin = icp_final_reduce_tagged(part, nact, seq, red, P.timeout_ns, &s_fail);
        if (threadIdx.x < 27) fin27[threadIdx.x] = fin;
        const unsigned long long ts3 = globaltimer_ns();
        ICP_BAR(); // fin27 and s_fail visible
        if (s_fail) return;
        if (blockIdx.x == 0)
        {
            // the iteration's result: sums + the pose they were computed with (posted writes, nobody waits for them)
            IcpHostSlot *slot = P.slots + k;
            if (threadIdx.x < 27) icp_post(slot->chunk, threadIdx.x, fin27[threadIdx.x], seq);
            else if (threadIdx.x >= 32 && threadIdx.x < 36)
            {
                const int c = threadIdx.x - 32;
                const float tagf = __uint_as_float((unsigned int)seq);
                const float4 q = c < 3 ? make_float4(spose[4 * c], spose[4 * c + 1], spose[4 * c + 2], tagf) : make_float4(spose[3], spose[7], spose[11], tagf);
                st_volatile_f4(slot->pose + 4 * c, q);
            }
            else if (threadIdx.x == 96)
            {
                volatile unsigned long long *pr = P.dbg->post_ns[k & 31]; // debug ring (kfb_debug_icp_ring)
                pr[0] = ts0; pr[1] = ts1; pr[2] = ts3; pr[3] = ts3;
            }
        }
        if (k + 1 >= P.total) return;
        if (threadIdx.x == 64) s_pred = icp_predict_pose(fin27, spose, npose) ? 1 : 0;
        ICP_BAR();
        if (!s_pred) return; // not positive definite: the same verdict in every CTA; the host takes over (ordinary launches)
        if (blockIdx.x == 0 && threadIdx.x == 0)
        {
            // CTAs that sat this level out pick the pose up here when their level starts
            if (P.lv[icp_level_of(P.iters, P.levels, k + 1)].nact > nact)
            {
                const unsigned int tag = (unsigned int)(seq + 1ull);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                {
                    if (c < 3) st_volatile_f4(P.devgate->chunk + 4 * c, gate_chunk(npose[4 * c], npose[4 * c + 1], npose[4 * c + 2], tag));
                    else st_volatile_f4(P.devgate->chunk + 12, gate_chunk(npose[3], npose[7], npose[11], tag));
                }
            }
            const unsigned long long ts4 = globaltimer_ns();
            P.dbg->post_ns[k & 31][3] = ts4;
        }
        ++k;
    }
}

// CTAs of an ICP grid: one per SM, and no more than the whole-schedule kernel's final sum keeps in registers per thread
// (ICP_FREE_MAXM partials x 15 slices) -- on a part with more SMs the surplus ones stay idle during ICP
static int icp_grid(const kfb_ctx *ctx) { return std::min(std::min(ctx->sm_count, (ICP_THREADS / 32) * ICP_FREE_MAXM), ctx->icp_tagged_cap); }

static int icp_setup(kfb_ctx *ctx, int level, IcpArgs &a, int &blocks)
{
    if (level < 0 || level >= ctx->levels) { ctx->err = "icp level out of range"; return KFB_ERR_INVALID; }
    const Level &L = ctx->L[level];
    a.cur_v = L.v[ctx->cur]; a.cur_n = L.n[ctx->cur];
    a.pre_v = L.v[ctx->prev]; a.pre_n = L.n[ctx->prev];
    a.k = L.k;
    if (ctx->p.compat_icp_rows) { a.cov_w = (L.k.w / 32) * 32; a.cov_h = (L.k.h / 32) * 32; }
    else { a.cov_w = L.k.w; a.cov_h = L.k.h; }
    a.dist_thres = ctx->p.icp_dist_threshold;
    a.sine_thres = sinf(ctx->p.icp_angle_threshold * 0.017453293f); // icp_registration.cpp:5, types.hpp:81
    a.partials = ctx->icp_partials;
    a.ticket = ctx->icp_ticket;
    a.out = ctx->icp_dev;
    // CTAs that take part at this level: one pixel per thread while the SMs last.  Measured on B200 (tools/icp_ab.py,
    // whole-schedule kernel per frame, round 1's protocol): KFB_ICP_PXT = 1 / 2 / 4 / 8 pixels per thread -> 184 / 189 / 209 / 282 us: every
    // further pixel per thread costs more in the accumulation than the smaller reduction saves.  Also measured and
    // dropped: the coarsest level inside ONE thread-block cluster of 8 CTAs (4 pixels per thread), partial sums and
    // next pose handed over through distributed shared memory with cluster-scope release / acquire instead of the
    // global partials + ticket + device gate: 9.9 us per iteration against 7.5 us (profiles/README.md).  The direct
    // and the whole-schedule kernel use the same count, hence the same pixel -> thread map and summation order.
    const int npix = a.cov_w * a.cov_h;
    int pxt = 1;
    if (const char *e = getenv("KFB_ICP_PXT")) { const int v = atoi(e); if (v >= 1 && v <= 64) pxt = v; }
    blocks = npix > 0 ? std::min(icp_grid(ctx), (npix + ICP_THREADS * pxt - 1) / (ICP_THREADS * pxt)) : 0;
    a.stride = blocks * ICP_THREADS;
    return KFB_OK;
}

#define ICP_RC_NO_RESULT (-7) // internal: the stream went idle without the awaited result (never returned through the C-ABI)
// spin on the tagged result chunks in mapped host memory (with a stream query as the failure detector)
static int icp_wait(kfb_ctx *ctx, unsigned long long seq, double out27[27])
{
    volatile IcpHostResult *h = ctx->icp_host;
    unsigned long spins = 0;
    for (;;)
    {
        int ready = 0;
        for (int i = 0; i < 27; ++i) ready += (h->chunk[i].tag == seq);
        if (ready == 27) break;
        if ((++spins & 0xfffff) == 0)
        {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) KFB_CUDA(ctx, q);
            if (q == cudaSuccess)
            {
                KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                ready = 0;
                for (int i = 0; i < 27; ++i) ready += (h->chunk[i].tag == seq);
                if (ready == 27) break;
                ctx->err = "icp result never arrived (gate timeout or abort)";
                return ICP_RC_NO_RESULT;
            }
        }
    }
    __sync_synchronize();
    // a chunk is written by one aligned 16-byte store, so value and tag arrive together
    for (int i = 0; i < 27; ++i) out27[i] = h->chunk[i].value;
    // everything enqueued before this ICP has run: a composite that gave up on a peer has said so by now
    return check_device_error(ctx);
}

int launch_icp(kfb_ctx *ctx, int level, const float pose12[12], double out27[27])
{
    IcpArgs a;
    int blocks = 0;
    const int rc = icp_setup(ctx, level, a, blocks);
    if (rc) return rc;
    a.pose = make_pose(pose12);
    a.seq = ++ctx->icp_seq;
    if (blocks <= 0)
    {
        for (int i = 0; i < 27; ++i) out27[i] = 0.0;
        return KFB_OK;
    }
    // a kernel that was torn down mid-reduction (timeout, abort) may have left the ticket non-zero
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->icp_ticket, 0, sizeof(unsigned int), ctx->stream));
    icp_kernel<<<blocks, ICP_THREADS, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    const int rcw = icp_wait(ctx, a.seq, out27);
    return rcw == ICP_RC_NO_RESULT ? KFB_ERR_CUDA : rcw; // an ordinary kernel that posts nothing is a device fault
}

// ---- whole-schedule driver -------------------------------------------------------------------------------
int icp_begin(kfb_ctx *ctx, const int *iters_per_level)
{
    IcpSchedule &S = ctx->icp_sched;
    if (S.active) { ctx->err = "icp schedule already active"; return KFB_ERR_INVALID; }
    S.total = 0;
    for (int l = 0; l < KFB_MAX_LEVELS; ++l)
    {
        S.iters[l] = l < ctx->levels ? iters_per_level[l] : 0;
        if (S.iters[l] < 0) S.iters[l] = 0;
        S.total += S.iters[l];
    }
    // one result slot per iteration in mapped host memory (KFB_ICP_MAX_ITERS)
    if (S.total > KFB_ICP_MAX_ITERS) { ctx->err = "icp schedule longer than 255 iterations"; return KFB_ERR_INVALID; }
    S.seq0 = ctx->icp_seq;
    S.enq = S.done = 0;
    S.direct = getenv("KFB_ICP_DIRECT") ? 1 : 0;
    // after a transport timeout (typically a profiler or debugger that serialises launches: the host cannot
    // answer a kernel whose launch call has not returned) the next schedules stay on ordinary launches
    if (ctx->icp_direct_left > 0) { --ctx->icp_direct_left; S.direct = 1; }
    S.active = 1;
    return KFB_OK;
}

// co-resident grid or nothing: cooperative launch (the driver checks), KFB_ICP_PLAIN_LAUNCH=1: ordinary launch
// behind an occupancy query.  cudaSuccess, or the reason the grid cannot be resident (the caller falls back).
template <class Args>
static cudaError_t icp_launch_resident(kfb_ctx *ctx, void (*kernel)(const Args), const Args &P, int blocks, int threads, size_t smem)
{
    if (getenv("KFB_ICP_PLAIN_LAUNCH"))
    {
        int per_sm = 0;
        cudaError_t le = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
        if (le == cudaSuccess && per_sm * ctx->sm_count < blocks) le = cudaErrorCooperativeLaunchTooLarge;
        if (le != cudaSuccess) return le;
        kernel<<<blocks, threads, smem, ctx->stream>>>(P);
        return cudaGetLastError();
    }
    void *kargs[] = {(void *)&P};
    return cudaLaunchCooperativeKernel((const void *)kernel, dim3(blocks), dim3(threads), kargs, smem, ctx->stream);
}
static bool icp_launch_refused(cudaError_t le)
{
    return le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources || le == cudaErrorNotSupported ||
           le == cudaErrorInvalidConfiguration;
}

static int icp_launch_freerun(kfb_ctx *ctx, const float pose12[12])
{
    IcpSchedule &S = ctx->icp_sched;
    IcpFreeArgs P;
    memset(&P, 0, sizeof(P));
    for (int l = 0; l < ctx->levels; ++l)
    {
        IcpArgs a;
        int blocks = 0;
        const int rc = icp_setup(ctx, l, a, blocks);
        if (rc) return rc;
        if (S.iters[l] > 0 && a.cov_w * a.cov_h <= 0) { ctx->err = "icp level has no pixels to visit"; return KFB_ERR_INVALID; }
        P.lv[l].cur_v = a.cur_v; P.lv[l].cur_n = a.cur_n; P.lv[l].pre_v = a.pre_v; P.lv[l].pre_n = a.pre_n;
        P.lv[l].k = a.k; P.lv[l].cov_w = a.cov_w; P.lv[l].cov_h = a.cov_h; P.lv[l].nact = blocks;
        P.iters[l] = S.iters[l];
        P.dist_thres = a.dist_thres; P.sine_thres = a.sine_thres;
    }
    P.levels = ctx->levels;
    P.total = S.total;
    P.tagged = ctx->icp_tagged;
    P.cap = ctx->icp_tagged_cap;
    P.slots = ctx->icp_slots_dev;
    P.devgate = ctx->icp_devgate;
    P.dbg = ctx->icp_dev;
    P.seq0 = S.seq0;
    P.timeout_ns = KFB_ICP_GATE_TIMEOUT_NS;
    if (const char *e = getenv("KFB_ICP_TIMEOUT_NS")) { const long long v = atoll(e); if (v > 0) P.timeout_ns = (unsigned long long)v; }
    memcpy(P.pose0, pose12, sizeof(P.pose0));
    const int blocks = icp_grid(ctx);
    if (ctx->profiling) cudaEventRecord(ctx->events[54], ctx->stream); // the frame's whole ICP: 54 .. 55
    const size_t cache_bytes = (size_t)ICP_CACHE_SLOTS * 2 * ICP_THREADS * sizeof(float4);
    if (!(ctx->icp_smem_set & 2)) // per device, hence per context
    {
        KFB_CUDA(ctx, cudaFuncSetAttribute(icp_freerun_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cache_bytes));
        ctx->icp_smem_set |= 2;
    }
    const cudaError_t le = icp_launch_resident(ctx, icp_freerun_kernel, P, blocks, ICP_THREADS, cache_bytes);
    if (le != cudaSuccess)
    {
        (void)cudaGetLastError(); // launch-configuration errors are not sticky
        if (icp_launch_refused(le))
        {
            S.direct = 1;
            ctx->icp_fallbacks++;
            return KFB_OK;
        }
        KFB_CUDA(ctx, le);
    }
    ctx->launches++;
    if (ctx->profiling) cudaEventRecord(ctx->events[55], ctx->stream);
    S.enq = S.total;
    return KFB_OK;
}

// the free-running kernel's result of iteration `it`: sums + the pose they belong to.  KFB_OK with *same_pose = 0:
// the device ran this iteration with another pose than the caller's.
static int icp_wait_slot(kfb_ctx *ctx, int it, unsigned long long seq, const float pose12[12], double out27[27], int *same_pose)
{
    volatile IcpHostSlot *h = ctx->icp_slots_host + it;
    const unsigned int tag32 = (unsigned int)seq;
    unsigned long spins = 0;
    auto complete = [&]() {
        int ready = 0;
        for (int i = 0; i < 27; ++i) ready += (h->chunk[i].tag == seq);
        for (int c = 0; c < 4; ++c)
        {
            unsigned int t;
            const float f = h->pose[4 * c + 3];
            memcpy(&t, &f, 4);
            ready += (t == tag32);
        }
        return ready == 31;
    };
    for (;;)
    {
        if (complete()) break;
        if ((++spins & 0xfffff) == 0)
        {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) KFB_CUDA(ctx, q);
            if (q == cudaSuccess)
            {
                KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (complete()) break;
                ctx->err = "icp result never arrived (the kernel gave up or had no prediction)";
                return ICP_RC_NO_RESULT;
            }
        }
    }
    __sync_synchronize();
    for (int i = 0; i < 27; ++i) out27[i] = h->chunk[i].value;
    float used[12];
    for (int r = 0; r < 3; ++r)
    {
        for (int c = 0; c < 3; ++c) used[4 * r + c] = h->pose[4 * r + c];
        used[4 * r + 3] = h->pose[12 + r];
    }
    *same_pose = memcmp(used, pose12, sizeof(used)) == 0;
    return check_device_error(ctx);
}

int icp_step(kfb_ctx *ctx, const float pose12[12], double out27[27])
{
    IcpSchedule &S = ctx->icp_sched;
    if (!S.active || S.done >= S.total) { ctx->err = "icp_step outside an active schedule"; return KFB_ERR_INVALID; }
    const unsigned long long seq = S.seq0 + (unsigned long long)S.done + 1ull;
    int rc;
    if (!S.direct)
    {
        if (S.done == 0)
        {
            rc = icp_launch_freerun(ctx, pose12); // may switch the schedule to direct launches
            if (rc) return rc;
        }
        if (!S.direct)
        {
            int same = 0;
            rc = icp_wait_slot(ctx, S.done, seq, pose12, out27, &same);
            if (rc == KFB_OK && same)
            {
                ++S.done;
                ctx->icp_seq = seq;
                return KFB_OK;
            }
            if (rc != KFB_OK && rc != ICP_RC_NO_RESULT) return rc;
            // The device's pose for this iteration is not the caller's (another solver than the one the kernel
            // replicates, sin / cos one ulp apart), or the kernel left without a result (a poll ran into its bound,
            // a system that is not positive definite).  Neither is a tracking failure: this and the remaining
            // iterations run as ordinary launches with the caller's poses, behind whatever the kernel still does.
            if (rc == KFB_OK) ctx->icp_mispredicts++;
            S.direct = 1;
            ctx->icp_fallbacks++;
            ctx->icp_direct_left = 64;
            const unsigned long long last = S.seq0 + (unsigned long long)S.total;
            if (last > ctx->icp_seq) ctx->icp_seq = last;
            ctx->err.clear();
        }
    }
    {
        // one ordinary launch per iteration: the fallback above, and KFB_ICP_DIRECT=1 for profilers that replay
        // kernels.  Same sums, bit for bit (see icp_setup).
        int k = S.done, level = ctx->levels - 1;
        while (level >= 0 && k >= S.iters[level]) { k -= S.iters[level]; --level; }
        rc = launch_icp(ctx, level, pose12, out27);
        if (rc) return rc;
        ++S.done;
        return KFB_OK;
    }
}

int icp_end(kfb_ctx *ctx)
{
    IcpSchedule &S = ctx->icp_sched;
    if (!S.active) return KFB_OK;
    // iterations never stepped (early exit / tracking failure): the kernel runs its schedule out by itself; its slots
    // carry sequence numbers no later schedule uses
    const unsigned long long last = S.seq0 + (unsigned long long)S.total;
    if (S.enq > 0) ctx->icp_seq = last > ctx->icp_seq ? last : ctx->icp_seq;
    S.active = 0;
    return KFB_OK;
}

} // namespace kfb
