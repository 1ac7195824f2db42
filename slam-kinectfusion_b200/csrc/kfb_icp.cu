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
// icp_kernel (one launch per kfb_icp_accumulate) and icp_persistent_kernel (the whole schedule in one launch,
// with the host round trip hidden by verified pose prediction -- see the comment at that kernel).
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

#define ICP_THREADS 480                 // worker threads of a CTA (15 warps); the persistent kernel adds one service warp
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

__device__ __forceinline__ unsigned long long ld_volatile_u64(const volatile unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
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
// frame's vertex / normal of a pixel do not change while the pose does: the persistent kernel keeps the first
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
__device__ __forceinline__ void icp_post(IcpHostResult *out, int i, double v, unsigned long long seq)
{
    asm volatile("st.volatile.global.v2.b64 [%0], {%1, %2};" ::"l"(&out->chunk[i]), "l"(__double_as_longlong(v)), "l"(seq) : "memory");
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
    if (threadIdx.x < 27) icp_post(a.out, threadIdx.x, fin, a.seq);
}

// ---- the whole coarse-to-fine loop in ONE persistent kernel ----------------------------------------------
// kfb_icp_begin/step/end: the grid (one CTA per SM, all co-resident) runs every iteration of the schedule.
// Per iteration: all CTAs accumulate their pixels and publish a partial; the last CTA to arrive (ticket)
// reduces and posts the 27 tagged sums to mapped host memory, where the host does the reference's 6x6 solve
// and publishes the next pose in its mapped gate.  The PCIe round trip of that exchange is taken off the
// critical path by SPECULATION: the last CTA also solves the system itself, with the host's operations in the
// host's order (IEEE double add/mul/div/sqrt are exactly rounded on both sides; only sin/cos may differ in the
// last bit, which survives the cast to float with probability ~2^-29), and releases the grid with that pose at
// once.  A service warp of CTA 0 -- the only PCIe reader -- mirrors each pose the host publishes into device
// memory; at the end of a speculative iteration the pose it used is compared, bit for bit, with the host's:
// equal => its sums are posted, different => the iteration is repeated with the host's pose.  The host's solve
// stays authoritative and every result equals the non-speculative schedule's (KFB_ICP_NOSPEC=1 runs that).
// Polls are bounded (KFB_ICP_GATE_TIMEOUT_NS, KFB_ICP_TIMEOUT_NS=<ns> overrides); on timeout or abort every CTA leaves and
// the host finishes the schedule with ordinary launches (icp_step).
#define ICP_MIRROR_RING 64
struct IcpMirror // device memory, written by the service warp
{
    unsigned long long tag[ICP_MIRROR_RING];   // seq of the iteration the pose is for; | 1<<63 = abort
    float pose[ICP_MIRROR_RING][12];
};
struct IcpPersistArgs
{
    IcpLevel lv[KFB_MAX_LEVELS];
    int iters[KFB_MAX_LEVELS];
    int levels, total;
    float dist_thres, sine_thres;
    double *partials;
    unsigned int *ticket;
    IcpHostResult *out;
    const IcpHostGate *gate;
    IcpDevGate *devgate;
    IcpMirror *mirror;
    unsigned long long seq0;   // iteration k carries sequence number seq0 + k + 1
    unsigned long long round0; // release counter base (monotonic across schedules)
    int speculate;
    unsigned long long timeout_ns; // bound of every poll (x1 host pose, x2 mirrored pose, x3 device gate)
    float pose0[12];
};

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

enum { ICP_CMD_RUN = 0, ICP_CMD_LEAVE = 1 };

__global__ void __launch_bounds__(ICP_THREADS + 32) icp_persistent_kernel(const IcpPersistArgs P)
{
    __shared__ double sm[ICP_THREADS / 32][27];
    __shared__ double red[ICP_THREADS / 32][28];
    __shared__ double fin27[27];
    __shared__ bool is_last;
    __shared__ int s_cmd, s_iter, s_spec, s_ok, s_pred;
    __shared__ float spose[12], npose[12], hpose[12];
    extern __shared__ float4 cur_cache[]; // [ICP_CACHE_SLOTS][normal, vertex][ICP_THREADS]

    // ---- service warp: CTA 0 mirrors the host's poses into device memory; elsewhere it has nothing to do -------
    if (threadIdx.x >= ICP_THREADS)
    {
        if (blockIdx.x != 0 || threadIdx.x != ICP_THREADS) return;
        for (int j = 1; j < P.total; ++j)
        {
            // Gate: four 16-byte chunks {3 pose floats, tag}; the host rewrites each chunk with one aligned
            // 16-byte store and the tag is the low 32 bits of the sequence number, so a chunk is either wholly
            // old or wholly new and one poll (4 loads in flight) yields a consistent pose.
            const unsigned long long want = P.seq0 + (unsigned long long)j + 1ull;
            const unsigned int tag = (unsigned int)want;
            const unsigned long long t0 = globaltimer_ns();
            bool ok = false;
            float4 c0, c1, c2, c3;
            for (;;)
            {
                c0 = ld_volatile_f4(P.gate->chunk);
                c1 = ld_volatile_f4(P.gate->chunk + 4);
                c2 = ld_volatile_f4(P.gate->chunk + 8);
                c3 = ld_volatile_f4(P.gate->chunk + 12);
                const unsigned long long ab = ld_volatile_u64(&P.gate->abort_upto);
                if (ab >= want) break;
                if (__float_as_uint(c0.w) == tag && __float_as_uint(c1.w) == tag && __float_as_uint(c2.w) == tag &&
                    __float_as_uint(c3.w) == tag) { ok = true; break; }
                if (globaltimer_ns() - t0 > P.timeout_ns) break;
            }
            if (!ok)
            {
                // abort / timeout: every pose still awaited is answered with "leave"
                for (int r = j; r < P.total; ++r)
                    *(volatile unsigned long long *)&P.mirror->tag[r % ICP_MIRROR_RING] = (P.seq0 + (unsigned long long)r + 1ull) | (1ull << 63);
                return;
            }
            volatile float *d = P.mirror->pose[j % ICP_MIRROR_RING]; // chunk r = {R[r][0..2]}, chunk 3 = t
            d[0] = c0.x; d[1] = c0.y; d[2] = c0.z; d[3] = c3.x;
            d[4] = c1.x; d[5] = c1.y; d[6] = c1.z; d[7] = c3.y;
            d[8] = c2.x; d[9] = c2.y; d[10] = c2.z; d[11] = c3.z;
            __threadfence();
            *(volatile unsigned long long *)&P.mirror->tag[j % ICP_MIRROR_RING] = want;
        }
        return;
    }

    // ---- workers --------------------------------------------------------------------------------------------
    IcpArgs a;
    a.dist_thres = P.dist_thres; a.sine_thres = P.sine_thres;
    if (threadIdx.x < 12) spose[threadIdx.x] = P.pose0[threadIdx.x];
    int k = 0, spec_used = 0;
    int cached_level = -1; // level whose current-frame pixels this thread holds in cur_cache
    unsigned long long round = P.round0;
    ICP_BAR();
    for (;;)
    {
        // level of iteration k (coarse to fine)
        int level = P.levels - 1, kk = k;
        while (level > 0 && kk >= P.iters[level]) { kk -= P.iters[level]; --level; }
        const IcpLevel &L = P.lv[level];
        a.cur_v = L.cur_v; a.cur_n = L.cur_n; a.pre_v = L.pre_v; a.pre_n = L.pre_n;
        a.k = L.k; a.cov_w = L.cov_w; a.cov_h = L.cov_h;
        const unsigned long long seq = P.seq0 + (unsigned long long)k + 1ull;
#pragma unroll
        for (int i = 0; i < 9; ++i) a.pose.R.m[i] = spose[4 * (i / 3) + (i % 3)];
#pragma unroll
        for (int i = 0; i < 3; ++i) a.pose.t[i] = spose[4 * i + 3];
        const unsigned long long ts0 = globaltimer_ns();
        // only the CTAs that own pixels at this level take part in the reduction
        const int nact = L.nact;
        unsigned long long ts1 = ts0;
        if (threadIdx.x == 0) is_last = false;
        if ((int)blockIdx.x < nact)
        {
            double acc[27];
#pragma unroll
            for (int i = 0; i < 27; ++i) acc[i] = 0.0;
            icp_accumulate_pixels<true>(a, acc, blockIdx.x * ICP_THREADS + threadIdx.x, nact * ICP_THREADS, cur_cache, level != cached_level);
            cached_level = level;
            ts1 = globaltimer_ns();
            const double s = icp_block_reduce(acc, sm);
            if (threadIdx.x < 27)
            {
                P.partials[(size_t)blockIdx.x * 27 + threadIdx.x] = s;
                __threadfence();
            }
            ICP_BAR();
            if (threadIdx.x == 0)
            {
                const unsigned int t = atomicInc(P.ticket, nact - 1);
                is_last = (t == (unsigned)(nact - 1));
            }
        }
        ICP_BAR();
        if (is_last)
        {
            const unsigned long long ts2 = globaltimer_ns();
            __threadfence();
            const double fin = icp_final_reduce(P.partials, nact, red);
            if (threadIdx.x < 27) fin27[threadIdx.x] = fin;
            const unsigned long long ts3 = globaltimer_ns();
            ICP_BAR(); // fin27 visible
            // (1) thread 0: a speculative iteration is only as good as its pose -- compare with the host's, bit for
            //     bit (hpose = the host's pose for this iteration).  Concurrently thread 32 (another warp) predicts
            //     the next pose from the sums; the prediction is only used if (1) passes.
            if (threadIdx.x == 0)
            {
                int ok = 1, cmd = ICP_CMD_RUN;
                if (spec_used)
                {
                    const int slot = k % ICP_MIRROR_RING;
                    const unsigned long long t0 = globaltimer_ns();
                    unsigned long long v;
                    for (;;)
                    {
                        v = ld_volatile_u64(&P.mirror->tag[slot]);
                        if ((v & ~(1ull << 63)) == seq) break;
                        if (globaltimer_ns() - t0 > 2ull * P.timeout_ns) { v = 1ull << 63; break; }
                    }
                    if (v >> 63) cmd = ICP_CMD_LEAVE;
                    else
                    {
                        for (int i = 0; i < 12; ++i)
                        {
                            const float h = __ldcg(&P.mirror->pose[slot][i]);
                            hpose[i] = h;
                            if (__float_as_uint(h) != __float_as_uint(spose[i])) ok = 0;
                        }
                    }
                }
                s_ok = ok; s_cmd = cmd;
            }
            if (threadIdx.x == 32) s_pred = (P.speculate && k + 1 < P.total) ? (icp_predict_pose(fin27, spose, npose) ? 1 : 0) : 0;
            ICP_BAR();
            if (s_cmd == ICP_CMD_RUN && s_ok)
            {
                if (threadIdx.x < 27) icp_post(P.out, threadIdx.x, fin27[threadIdx.x], seq);
            }
            // (2) decide what the grid does next
            if (threadIdx.x == 0)
            {
                const unsigned long long ts4 = globaltimer_ns();
                const bool posted = s_cmd == ICP_CMD_RUN && s_ok;
                const int k_posted = k;
                int cmd = s_cmd, iter = k, spec = 0;
                const float *next = npose;
                if (cmd == ICP_CMD_RUN)
                {
                    if (!s_ok)
                    {
                        iter = k; spec = 0; next = hpose;                 // repeat iteration k with the host's pose
                        ((volatile unsigned long long *)P.out->stamps)[7] += 1ull; // debug: mispredictions since creation
                    }
                    else if (k + 1 == P.total) cmd = ICP_CMD_LEAVE;       // schedule complete
                    else
                    {
                        iter = k + 1;
                        spec = s_pred;
                        if (!spec)
                        {
                            // no prediction: wait for the host's pose (mirrored by the service warp)
                            const int slot = iter % ICP_MIRROR_RING;
                            const unsigned long long want = seq + 1ull, t0 = globaltimer_ns();
                            unsigned long long v;
                            for (;;)
                            {
                                v = ld_volatile_u64(&P.mirror->tag[slot]);
                                if ((v & ~(1ull << 63)) == want) break;
                                if (globaltimer_ns() - t0 > 2ull * P.timeout_ns) { v = 1ull << 63; break; }
                            }
                            if (v >> 63) cmd = ICP_CMD_LEAVE;
                            else
                            {
                                for (int i = 0; i < 12; ++i) hpose[i] = __ldcg(&P.mirror->pose[slot][i]);
                                next = hpose;
                            }
                        }
                    }
                }
                const unsigned long long ts5 = globaltimer_ns();
                {
                    // release: four self-validating chunks (see IcpDevGate).  Nothing else has to be ordered before
                    // them: the partials were consumed above, the ticket has wrapped to zero by itself.
                    const unsigned int tag = ((unsigned int)((round + 1ull) & 0xfffffull) << 12) | ((unsigned int)(iter & 0xff) << 4) |
                                             ((unsigned int)(spec & 1) << 2) | (unsigned int)(cmd & 3);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                    {
                        if (c < 3) st_volatile_f4(P.devgate->chunk + 4 * c, make_float4(next[4 * c], next[4 * c + 1], next[4 * c + 2], __uint_as_float(tag)));
                        else st_volatile_f4(P.devgate->chunk + 12, make_float4(next[3], next[7], next[11], __uint_as_float(tag)));
                    }
                }
                // debug hooks (kfb_debug_icp_stamps / _ring): host-memory stores, after the release on purpose
                const unsigned long long ts6 = globaltimer_ns();
                volatile unsigned long long *st = P.out->stamps;
                if (posted)
                {
                    st[0] = ts0; st[1] = ts1; st[2] = ts2; st[3] = ts3; st[4] = ts4;
                    volatile unsigned long long *pr = P.out->post_ns[k_posted & 31];
                    pr[0] = ts0; pr[1] = ts3; pr[2] = ts4; pr[3] = ts6;
                }
                st[5] = ts5; st[6] = ts6;
            }
        }
        // every CTA (the last one included) picks its orders up from the device gate
        if (threadIdx.x == 0)
        {
            const unsigned int want = (unsigned int)((round + 1ull) & 0xfffffull);
            const unsigned long long t0 = globaltimer_ns();
            int cmd = ICP_CMD_RUN;
            float4 c0, c1, c2, c3;
            for (;;)
            {
                // one load per poll (the chunk written last) keeps the pollers out of the writer's way; CTAs without
                // pixels at this level are in no hurry at all
                c3 = ld_volatile_f4(P.devgate->chunk + 12);
                if ((__float_as_uint(c3.w) >> 12) == want)
                {
                    c0 = ld_volatile_f4(P.devgate->chunk);
                    c1 = ld_volatile_f4(P.devgate->chunk + 4);
                    c2 = ld_volatile_f4(P.devgate->chunk + 8);
                    const unsigned int t = __float_as_uint(c3.w);
                    if (__float_as_uint(c0.w) == t && __float_as_uint(c1.w) == t && __float_as_uint(c2.w) == t) break;
                }
                else if ((int)blockIdx.x >= nact) __nanosleep(200);
                if (globaltimer_ns() - t0 > 3ull * P.timeout_ns) { cmd = ICP_CMD_LEAVE; break; }
            }
            if (cmd == ICP_CMD_RUN)
            {
                const unsigned int t = __float_as_uint(c0.w);
                cmd = (int)(t & 3u);
                s_iter = (int)((t >> 4) & 0xffu);
                s_spec = (int)((t >> 2) & 1u);
                // rows 0..2 without the translation, then the translation column
                spose[0] = c0.x; spose[1] = c0.y; spose[2] = c0.z;
                spose[4] = c1.x; spose[5] = c1.y; spose[6] = c1.z;
                spose[8] = c2.x; spose[9] = c2.y; spose[10] = c2.z;
                spose[3] = c3.x; spose[7] = c3.y; spose[11] = c3.z;
            }
            s_cmd = cmd;
        }
        ICP_BAR();
        if (s_cmd != ICP_CMD_RUN) return;
        k = s_iter; spec_used = s_spec;
        ++round;
    }
}

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
    // persistent kernel per frame): KFB_ICP_PXT = 1 / 2 / 4 / 8 pixels per thread -> 184 / 189 / 209 / 282 us: every
    // further pixel per thread costs more in the accumulation than the smaller reduction saves.  Also measured and
    // dropped: the coarsest level inside ONE thread-block cluster of 8 CTAs (4 pixels per thread), partial sums and
    // next pose handed over through distributed shared memory with cluster-scope release / acquire instead of the
    // global partials + ticket + device gate: 9.9 us per iteration against 7.5 us (profiles/README.md).  The direct
    // and the persistent kernel use the same count, hence the same pixel -> thread map and summation order.
    const int npix = a.cov_w * a.cov_h;
    int pxt = 1;
    if (const char *e = getenv("KFB_ICP_PXT")) { const int v = atoi(e); if (v >= 1 && v <= 64) pxt = v; }
    blocks = npix > 0 ? std::min(ctx->sm_count, (npix + ICP_THREADS * pxt - 1) / (ICP_THREADS * pxt)) : 0;
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

// ---- persistent schedule -------------------------------------------------------------------------------
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
    // the release tag of the persistent kernel carries the iteration in 8 bits (IcpDevGate)
    if (S.total > 255) { ctx->err = "icp schedule longer than 255 iterations"; return KFB_ERR_INVALID; }
    S.seq0 = ctx->icp_seq;
    S.enq = S.done = 0;
    S.direct = getenv("KFB_ICP_DIRECT") ? 1 : 0;
    // after a transport timeout (typically a profiler or debugger that serialises launches: the host cannot
    // answer a kernel whose launch call has not returned) the next schedules stay on ordinary launches
    if (ctx->icp_direct_left > 0) { --ctx->icp_direct_left; S.direct = 1; }
    S.active = 1;
    return KFB_OK;
}

static int icp_launch_persistent(kfb_ctx *ctx, const float pose12[12])
{
    IcpSchedule &S = ctx->icp_sched;
    IcpPersistArgs P;
    memset(&P, 0, sizeof(P));
    int max_pix = 0;
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
        if (S.iters[l] > 0 && a.cov_w * a.cov_h > max_pix) max_pix = a.cov_w * a.cov_h;
    }
    P.levels = ctx->levels;
    P.total = S.total;
    P.partials = ctx->icp_partials;
    P.ticket = ctx->icp_ticket;
    P.out = ctx->icp_dev;
    P.gate = ctx->icp_gate_dev;
    P.devgate = ctx->icp_devgate;
    P.mirror = (IcpMirror *)ctx->icp_mirror;
    P.seq0 = S.seq0;
    P.round0 = ctx->icp_round;
    ctx->icp_round += (unsigned long long)(4 * S.total + 8); // rounds this launch can consume (repeats included)
    P.speculate = getenv("KFB_ICP_NOSPEC") ? 0 : 1;
    P.timeout_ns = KFB_ICP_GATE_TIMEOUT_NS;
    if (const char *e = getenv("KFB_ICP_TIMEOUT_NS")) { const long long v = atoll(e); if (v > 0) P.timeout_ns = (unsigned long long)v; }
    memcpy(P.pose0, pose12, sizeof(P.pose0));
    // every CTA must be resident at once (they wait on each other): one per SM (see icp_setup)
    const int blocks = ctx->sm_count;
    (void)max_pix;
    if (ctx->profiling) cudaEventRecord(ctx->events[54], ctx->stream); // the frame's whole ICP: 54 .. 55
    const size_t cache_bytes = (size_t)ICP_CACHE_SLOTS * 2 * ICP_THREADS * sizeof(float4);
    if (!ctx->icp_smem_set) // per device, hence per context
    {
        KFB_CUDA(ctx, cudaFuncSetAttribute(icp_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cache_bytes));
        ctx->icp_smem_set = 1;
    }
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->icp_ticket, 0, sizeof(unsigned int), ctx->stream));
    // The CTAs wait on each other through the device gate, so the whole grid has to be co-resident.  A
    // cooperative launch makes the driver check exactly that (SM limits under MPS / MIG included) and refuse
    // the launch otherwise; the caller then runs the schedule with one ordinary launch per iteration, which
    // gives the same sums bit for bit.  KFB_ICP_PLAIN_LAUNCH=1: ordinary launch behind an occupancy query.
    cudaError_t le;
    if (getenv("KFB_ICP_PLAIN_LAUNCH"))
    {
        int per_sm = 0;
        le = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_persistent_kernel, ICP_THREADS + 32, cache_bytes);
        if (le == cudaSuccess && per_sm * ctx->sm_count < blocks) le = cudaErrorCooperativeLaunchTooLarge;
        if (le == cudaSuccess)
        {
            icp_persistent_kernel<<<blocks, ICP_THREADS + 32, cache_bytes, ctx->stream>>>(P);
            le = cudaGetLastError();
        }
    }
    else
    {
        void *kargs[] = {(void *)&P};
        le = cudaLaunchCooperativeKernel((const void *)icp_persistent_kernel, dim3(blocks), dim3(ICP_THREADS + 32), kargs, cache_bytes, ctx->stream);
    }
    if (le != cudaSuccess)
    {
        (void)cudaGetLastError(); // launch-configuration errors are not sticky
        if (le == cudaErrorCooperativeLaunchTooLarge || le == cudaErrorLaunchOutOfResources || le == cudaErrorNotSupported ||
            le == cudaErrorInvalidConfiguration)
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

int icp_step(kfb_ctx *ctx, const float pose12[12], double out27[27])
{
    IcpSchedule &S = ctx->icp_sched;
    if (!S.active || S.done >= S.total) { ctx->err = "icp_step outside an active schedule"; return KFB_ERR_INVALID; }
    const unsigned long long seq = S.seq0 + (unsigned long long)S.done + 1ull;
    int rc;
    if (S.done == 0 && !S.direct)
    {
        rc = icp_launch_persistent(ctx, pose12); // first pose by parameter; may switch the schedule to direct launches
        if (rc) return rc;
    }
    else if (!S.direct)
    {
        // publish this iteration's pose: four tagged 16-byte chunks (x86 TSO keeps each store whole);
        // the last CTA of the previous iteration is polling for it
        IcpHostGate *g = ctx->icp_gate_host;
        const unsigned int tag = (unsigned int)seq;
        float tagf;
        memcpy(&tagf, &tag, 4);
        const __m128 k0 = _mm_set_ps(tagf, pose12[2], pose12[1], pose12[0]);
        const __m128 k1 = _mm_set_ps(tagf, pose12[6], pose12[5], pose12[4]);
        const __m128 k2 = _mm_set_ps(tagf, pose12[10], pose12[9], pose12[8]);
        const __m128 k3 = _mm_set_ps(tagf, pose12[11], pose12[7], pose12[3]);
        _mm_store_ps((float *)g->chunk, k0);
        _mm_store_ps((float *)g->chunk + 4, k1);
        _mm_store_ps((float *)g->chunk + 8, k2);
        _mm_store_ps((float *)g->chunk + 12, k3);
        _mm_sfence();
    }
    if (!S.direct)
    {
        rc = icp_wait(ctx, seq, out27);
        if (rc == KFB_OK)
        {
            ++S.done;
            ctx->icp_seq = seq;
            return KFB_OK;
        }
        if (rc != ICP_RC_NO_RESULT) return rc;
        // Transport failure: the persistent kernel gave up waiting (a descheduled host thread, a debugger, a
        // profiler replaying kernels) and has left the device -- icp_wait saw the stream idle.  That is not a
        // tracking failure: the remaining iterations of this frame run as ordinary launches.
        S.direct = 1;
        S.enq = 0;
        ctx->icp_fallbacks++;
        ctx->icp_direct_left = 64;
        const unsigned long long last = S.seq0 + (unsigned long long)S.total;
        if (last > ctx->icp_seq) ctx->icp_seq = last; // no tag the dead kernel may have posted can be mistaken for a new one
        ctx->err.clear();
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
    // iterations never fed a pose (early exit / tracking failure) retire through the abort gate: the polling
    // thread releases every CTA with the leave bit
    const unsigned long long last = S.seq0 + (unsigned long long)S.total;
    if (S.enq > 0 && S.done < S.total)
    {
        volatile IcpHostGate *g = ctx->icp_gate_host;
        g->abort_upto = last;
        __sync_synchronize();
    }
    if (S.enq > 0) ctx->icp_seq = last > ctx->icp_seq ? last : ctx->icp_seq;
    S.active = 0;
    return KFB_OK;
}

} // namespace kfb
