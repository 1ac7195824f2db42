// kfb_icp.cu -- projective-data-association point-to-plane ICP, one fused kernel per
// iteration (replaces kf::device::rigidICP, kfusion/src/rigid_icp.cu:46-169).
//
// One launch does correspondence search, gating, the residual row [s x n, n, n.(d-s)] and the
// 27 unique products of the 6x7 normal equations, reduces them with warp-shuffle trees, one
// shared-memory stage per block, and a single-pass grid reduction (last-block-done ticket,
// fixed summation order => bit-reproducible), and writes the 27 sums straight into mapped
// pinned host memory followed by a sequence flag.  The host spins on the flag: no cudaMemcpy,
// no allocation, no second kernel, no device sync per iteration (the reference does
// 2 launches + 2 cudaMalloc + 2 cudaFree + a blocking memcpy).
//
// Numerics: products are f32 exactly as in the reference
// (`smem[tid] = row[i]*row[j]`), sums are carried in f64 end to end (the reference rounds
// per-32x32-tile sums to f32 in between; the difference is ~1e-7 relative per entry and far
// below the 1e-4 pose tolerance).  Coverage: compat_icp_rows reproduces the reference's
// truncated grid floor(w/32) x floor(h/32) tiles (§9 Q7).
#include "kfb_common.cuh"
#include <xmmintrin.h>
#include <cstring>

namespace kfb
{

struct IcpArgs
{
    const float4 *cur_v, *cur_n, *pre_v, *pre_n;
    Intr k;
    int cov_w, cov_h; // pixels visited: [0,cov_w) x [0,cov_h)
    Pose pose;
    float dist_thres, sine_thres;
    double *partials;           // [gridDim.x][27]
    unsigned int *ticket;
    IcpHostResult *out;         // mapped host memory
    unsigned long long seq;
    const IcpHostGate *gate;    // mapped host memory: pose published by the host (gated schedule)
    IcpDevGate *devgate;        // device memory: pose handed from one gated launch to the next
    int poll_next;              // this launch's tail fetches the next pose from the host gate
};

#define ICP_THREADS 256

// findCoresp (rigid_icp.cu:46-80) + row (rigid_icp.cu:85-95)
__device__ __forceinline__ bool icp_row(const IcpArgs &a, int x, int y, float row[7])
{
    const int i = y * a.k.w + x;
    const float4 nc4 = __ldg(a.cur_n + i);
    if (isnan(nc4.x)) return false;
    const float4 vc4 = __ldg(a.cur_v + i);
    const float3 r = rot3(a.pose.R, vc4.x, vc4.y, vc4.z);
    const float sx = __fadd_rn(r.x, a.pose.t[0]), sy = __fadd_rn(r.y, a.pose.t[1]), sz = __fadd_rn(r.z, a.pose.t[2]);
    // Intrs::proj (device_utils.cuh:15-21)
    const float qx = __fdividef(sx, sz), qy = __fdividef(sy, sz);
    const int px = __float2int_rn(__fmaf_rn(qx, a.k.fx, a.k.cx));
    const int py = __float2int_rn(__fmaf_rn(qy, a.k.fy, a.k.cy));
    if (!(sz > 0.f && px >= 0 && py >= 0 && px < a.k.w && py < a.k.h)) return false;
    const int j = py * a.k.w + px;
    const float4 vp = __ldg(a.pre_v + j);
    const float dx = __fsub_rn(sx, vp.x), dy = __fsub_rn(sy, vp.y), dz = __fsub_rn(sz, vp.z);
    const float dist = __fsqrt_rn(dot3c(dx, dy, dz, dx, dy, dz));
    if (!(dist <= a.dist_thres)) return false;
    const float3 nc = rot3(a.pose.R, nc4.x, nc4.y, nc4.z);
    const float4 np = __ldg(a.pre_n + j);
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
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// GATED launches were enqueued ahead of time (kfb_icp_begin/step): their pose is not a launch
// parameter but sits in device memory (IcpDevGate), put there by the tail of the previous launch.
// That tail -- ONE thread of the last block, after it has posted the 27 sums -- polls the host's
// mapped gate for the next pose (or an abort), so exactly one PCIe reader exists at any time and
// the launch latency of iteration k+1 overlaps iteration k and the host's 6x6 solve.  The poll is
// bounded (KFB_ICP_GATE_TIMEOUT_NS); on timeout or abort the device gate is invalidated and every
// later gated launch returns at once.
template <bool GATED>
__global__ void __launch_bounds__(ICP_THREADS) icp_kernel(const IcpArgs a0)
{
    IcpArgs a = a0;
    if (GATED)
    {
        if (a.devgate->seq != a.seq) return; // uniform: written before this launch started
#pragma unroll
        for (int i = 0; i < 9; ++i) a.pose.R.m[i] = a.devgate->pose[4 * (i / 3) + (i % 3)];
#pragma unroll
        for (int i = 0; i < 3; ++i) a.pose.t[i] = a.devgate->pose[4 * i + 3];
    }
    const unsigned long long ts0 = globaltimer_ns();
    double acc[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) acc[i] = 0.0;

    const int npix = a.cov_w * a.cov_h;
    for (int p = blockIdx.x * ICP_THREADS + threadIdx.x; p < npix; p += gridDim.x * ICP_THREADS)
    {
        const int y = p / a.cov_w, x = p - y * a.cov_w;
        float row[7];
        if (icp_row(a, x, y, row))
        {
            int s = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 7; ++j) acc[s++] += (double)__fmul_rn(row[i], row[j]);
        }
    }
    const unsigned long long ts1 = globaltimer_ns();
    // warp tree
#pragma unroll
    for (int i = 0; i < 27; ++i)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
    }
    __shared__ double sm[ICP_THREADS / 32][27];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
    {
#pragma unroll
        for (int i = 0; i < 27; ++i) sm[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x < 27)
    {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < ICP_THREADS / 32; ++w) s += sm[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * 27 + threadIdx.x] = s;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        const unsigned int t = atomicInc(a.ticket, gridDim.x - 1); // wraps to 0 on the last block
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    const unsigned long long ts2 = globaltimer_ns();
    __threadfence();
    // last block: fixed-order sum of the per-block partials.  Thread t owns value v = t & 31 over the
    // block slice {t >> 5, t >> 5 + 8, ...}: all loads of a thread are independent (one L2 round trip),
    // the summation order is a fixed function of the grid size => bit-reproducible.
    __shared__ double red[ICP_THREADS / 32][28];
    {
        const int v = threadIdx.x & 31, slice = threadIdx.x >> 5;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        if (v < 27)
        {
            const int nb = (int)gridDim.x, stride = ICP_THREADS / 32;
            int bb = slice;
            for (; bb + 3 * stride < nb; bb += 4 * stride)
            {
                const double p0 = __ldcg(a.partials + (size_t)bb * 27 + v);
                const double p1 = __ldcg(a.partials + (size_t)(bb + stride) * 27 + v);
                const double p2 = __ldcg(a.partials + (size_t)(bb + 2 * stride) * 27 + v);
                const double p3 = __ldcg(a.partials + (size_t)(bb + 3 * stride) * 27 + v);
                s0 += p0; s1 += p1; s2 += p2; s3 += p3;
            }
            for (; bb < nb; bb += stride) s0 += __ldcg(a.partials + (size_t)bb * 27 + v);
            red[slice][v] = (s0 + s1) + (s2 + s3);
        }
    }
    __syncthreads();
    const unsigned long long ts3 = globaltimer_ns();
    // one thread posts the 27 sums to mapped host memory, one system fence, then the flag
    if (warp == 0)
    {
        double fin = 0.0;
        if (lane < 27)
        {
#pragma unroll
            for (int w = 0; w < ICP_THREADS / 32; ++w) fin += red[w][lane];
        }
#pragma unroll
        for (int i = 0; i < 27; ++i)
        {
            const double v = __shfl_sync(0xffffffffu, fin, i);
            if (lane == 0) a.out->sums[i] = v;
        }
        if (lane == 0)
        {
            const unsigned long long ts4 = globaltimer_ns();
            __threadfence_system();
            a.out->seq = a.seq;
            const unsigned long long ts5 = globaltimer_ns();
            volatile unsigned long long *st = a.out->stamps[a.seq & 31];
            st[0] = ts0; st[1] = ts1; st[2] = ts2; st[3] = ts3; st[4] = ts4; st[5] = ts5; st[6] = 0; st[7] = a.seq;
            if (a.poll_next)
            {
                // Gate: four 16-byte chunks {3 pose floats, tag}; the host rewrites each chunk with one
                // aligned 16-byte store and the tag is the low 32 bits of the sequence number, so a chunk is
                // either wholly old or wholly new and one poll (4 loads in flight) yields a consistent pose.
                const unsigned long long want = a.seq + 1ull;
                const unsigned int tag = (unsigned int)want;
                const unsigned long long t0 = globaltimer_ns();
                bool ok = false;
                float4 c0, c1, c2, c3;
                for (;;)
                {
                    c0 = ld_volatile_f4(a.gate->chunk);
                    c1 = ld_volatile_f4(a.gate->chunk + 4);
                    c2 = ld_volatile_f4(a.gate->chunk + 8);
                    c3 = ld_volatile_f4(a.gate->chunk + 12);
                    const unsigned long long ab = ld_volatile_u64(&a.gate->abort_upto);
                    if (ab >= want) break;
                    if (__float_as_uint(c0.w) == tag && __float_as_uint(c1.w) == tag && __float_as_uint(c2.w) == tag &&
                        __float_as_uint(c3.w) == tag) { ok = true; break; }
                    if (globaltimer_ns() - t0 > KFB_ICP_GATE_TIMEOUT_NS) break;
                }
                if (ok)
                {
                    float *d = a.devgate->pose; // chunk r = {R[r][0..2]}, chunk 3 = t
                    d[0] = c0.x; d[1] = c0.y; d[2] = c0.z; d[3] = c3.x;
                    d[4] = c1.x; d[5] = c1.y; d[6] = c1.z; d[7] = c3.y;
                    d[8] = c2.x; d[9] = c2.y; d[10] = c2.z; d[11] = c3.z;
                }
                a.devgate->seq = ok ? want : 0ull;
                st[6] = globaltimer_ns();
            }
        }
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
    a.gate = ctx->icp_gate_dev;
    a.devgate = ctx->icp_devgate;
    a.poll_next = 0;
    const int npix = a.cov_w * a.cov_h;
    // latency-bound at the coarse levels: one pixel per thread until the grid covers 2 CTAs/SM
    blocks = (npix + ICP_THREADS - 1) / ICP_THREADS;
    if (blocks > 296) blocks = 296;
    return KFB_OK;
}

// spin on the mapped result flag (with a stream query as the failure detector)
static int icp_wait(kfb_ctx *ctx, unsigned long long seq, double out27[27])
{
    IcpHostResult *h = ctx->icp_host;
    unsigned long spins = 0;
    while (h->seq != seq)
    {
        if ((++spins & 0xfffff) == 0)
        {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) KFB_CUDA(ctx, q);
            if (q == cudaSuccess && h->seq != seq)
            {
                KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (h->seq != seq) { ctx->err = "icp result flag never arrived"; return KFB_ERR_CUDA; }
            }
        }
    }
    __sync_synchronize();
    for (int i = 0; i < 27; ++i) out27[i] = h->sums[i];
    return KFB_OK;
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
    icp_kernel<false><<<blocks, ICP_THREADS, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    return icp_wait(ctx, a.seq, out27);
}

// ---- gated, pre-enqueued schedule ---------------------------------------------------------------------
// Launch k of the schedule carries sequence number seq0 + k + 1.  Launch 0 takes its pose as a
// parameter (it is launched by the first kfb_icp_step); every launch but the last polls the host gate
// for its successor's pose in its tail.
static int icp_enqueue_next(kfb_ctx *ctx, const float *pose12_first)
{
    IcpSchedule &S = ctx->icp_sched;
    if (S.enq >= S.total) return KFB_OK;
    int k = S.enq, level = ctx->levels - 1; // flat index -> (level, iteration), coarse to fine
    while (level >= 0 && k >= S.iters[level]) { k -= S.iters[level]; --level; }
    IcpArgs a;
    int blocks = 0;
    const int rc = icp_setup(ctx, level, a, blocks);
    if (rc) return rc;
    if (blocks <= 0) { ctx->err = "icp level has no pixels to visit"; return KFB_ERR_INVALID; }
    a.seq = S.seq0 + (unsigned long long)S.enq + 1ull;
    a.poll_next = (S.enq + 1 < S.total) ? 1 : 0;
    if (pose12_first)
    {
        a.pose = make_pose(pose12_first);
        icp_kernel<false><<<blocks, ICP_THREADS, 0, ctx->stream>>>(a);
    }
    else
    {
        a.pose = make_pose(S.identity);
        icp_kernel<true><<<blocks, ICP_THREADS, 0, ctx->stream>>>(a);
    }
    KFB_LAUNCH_CHECK(ctx);
    ++S.enq;
    return KFB_OK;
}

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
    for (int i = 0; i < 12; ++i) S.identity[i] = (i % 5 == 0) ? 1.f : 0.f;
    S.seq0 = ctx->icp_seq;
    S.enq = S.done = 0;
    S.active = 1;
    return KFB_OK;
}

int icp_step(kfb_ctx *ctx, const float pose12[12], double out27[27])
{
    IcpSchedule &S = ctx->icp_sched;
    if (!S.active || S.done >= S.total) { ctx->err = "icp_step outside an active schedule"; return KFB_ERR_INVALID; }
    const unsigned long long seq = S.seq0 + (unsigned long long)S.done + 1ull;
    int rc;
    if (S.done == 0)
    {
        rc = icp_enqueue_next(ctx, pose12); // head of the chain: pose by parameter
        if (rc) return rc;
        for (int d = 0; d < KFB_ICP_LOOKAHEAD; ++d)
            if ((rc = icp_enqueue_next(ctx, nullptr)) != KFB_OK) return rc;
    }
    else
    {
        // publish this iteration's pose: payload first, then the flag (x86 TSO + compiler barriers);
        // the tail of the previous launch is polling for it
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
        if ((rc = icp_enqueue_next(ctx, nullptr)) != KFB_OK) return rc; // keep the queue KFB_ICP_LOOKAHEAD deep
    }
    rc = icp_wait(ctx, seq, out27);
    if (rc) return rc;
    ++S.done;
    ctx->icp_seq = seq;
    return KFB_OK;
}

int icp_end(kfb_ctx *ctx)
{
    IcpSchedule &S = ctx->icp_sched;
    if (!S.active) return KFB_OK;
    // launches enqueued but never fed a pose (early exit / tracking failure) retire through the abort
    // gate: the polling tail invalidates the device gate and the rest return at once
    const unsigned long long last = S.seq0 + (unsigned long long)S.enq;
    if (S.done < S.enq)
    {
        volatile IcpHostGate *g = ctx->icp_gate_host;
        g->abort_upto = last;
        __sync_synchronize();
    }
    ctx->icp_seq = last > ctx->icp_seq ? last : ctx->icp_seq;
    S.active = 0;
    return KFB_OK;
}

} // namespace kfb
