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

__global__ void __launch_bounds__(ICP_THREADS) icp_kernel(const IcpArgs a)
{
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
    __threadfence();
    // last block: fixed-order sum of the per-block partials, 8 warps x 32 lanes over blocks
    for (int v = warp; v < 27; v += ICP_THREADS / 32)
    {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(a.partials + (size_t)b * 27 + v);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0)
        {
            a.out->sums[v] = s;
            __threadfence_system();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence_system();
        a.out->seq = a.seq;
    }
}

int launch_icp(kfb_ctx *ctx, int level, const float pose12[12], double out27[27])
{
    if (level < 0 || level >= ctx->levels) { ctx->err = "icp level out of range"; return KFB_ERR_INVALID; }
    const Level &L = ctx->L[level];
    IcpArgs a;
    a.cur_v = L.v[ctx->cur]; a.cur_n = L.n[ctx->cur];
    a.pre_v = L.v[ctx->prev]; a.pre_n = L.n[ctx->prev];
    a.k = L.k;
    if (ctx->p.compat_icp_rows) { a.cov_w = (L.k.w / 32) * 32; a.cov_h = (L.k.h / 32) * 32; }
    else { a.cov_w = L.k.w; a.cov_h = L.k.h; }
    a.pose = make_pose(pose12);
    a.dist_thres = ctx->p.icp_dist_threshold;
    a.sine_thres = sinf(ctx->p.icp_angle_threshold * 0.017453293f); // icp_registration.cpp:5, types.hpp:81
    a.partials = ctx->icp_partials;
    a.ticket = ctx->icp_ticket;
    a.out = ctx->icp_dev;
    a.seq = ++ctx->icp_seq;
    const int npix = a.cov_w * a.cov_h;
    if (npix <= 0)
    {
        for (int i = 0; i < 27; ++i) out27[i] = 0.0;
        return KFB_OK;
    }
    // latency-bound at the coarse levels: one pixel per thread until the grid covers 2 CTAs/SM
    int blocks = (npix + ICP_THREADS - 1) / ICP_THREADS;
    if (blocks > 296) blocks = 296;
    icp_kernel<<<blocks, ICP_THREADS, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    // spin on the mapped flag (with a stream query as the failure detector)
    IcpHostResult *h = ctx->icp_host;
    unsigned long spins = 0;
    while (h->seq != a.seq)
    {
        if ((++spins & 0xfffff) == 0)
        {
            cudaError_t q = cudaStreamQuery(ctx->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) KFB_CUDA(ctx, q);
            if (q == cudaSuccess && h->seq != a.seq)
            {
                // kernel finished; the posted write must be visible after a sync
                KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                if (h->seq != a.seq) { ctx->err = "icp result flag never arrived"; return KFB_ERR_CUDA; }
            }
        }
    }
    __sync_synchronize();
    for (int i = 0; i < 27; ++i) out27[i] = h->sums[i];
    return KFB_OK;
}

} // namespace kfb
