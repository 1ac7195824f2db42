// kfb_frontend.cu -- depth front end: replaces cv::cuda::pyrDown, cv::cuda::bilateralFilter,
// kf::device::depthTruncation / getVertexmap / getNormalmap (kfusion/src/kinectfusion.cpp:48-76,
// kfusion/src/image_process.cu:8-94) and kf::device::resizePointsNormals (image_process.cu:95-135).
//
// Launch plan per frame (4 launches, every pixel of every output is written, so the
// reference's 20 GpuMat::setTo(0) per frame disappear):
//   pyrdown_kernel   x(L-1)  raw mm depth, 5x5 Gaussian, REFLECT_101
//   bilateral_kernel x1      all levels in one grid (blockIdx.z = level): 13-tap bilateral on the
//                            smem-tiled raw level image, then mm->m and the 5 m cut (§10.2, §9 Q2-Q4)
//   vertex_normal_kernel x1  all levels: vertex map + central-difference normal from a smem tile
//                            of filtered depth (+1 halo); vertices of the 4 neighbours are recomputed
//                            from depth instead of re-read (§9 Q5, Q6)
#include "kfb_common.cuh"

namespace kfb
{

__device__ __forceinline__ int reflect101(int p, int len)
{
    // BORDER_REFLECT_101, valid for |overshoot| < len (radius 2, len >= 3)
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
    if (p < 0) p = 0; // degenerate tiny images
    return p;
}

// ---- pyrDown ------------------------------------------------------------------------------
// vertical 5-tap chain first, then horizontal, every `sum + w*v` one FMA (upstream pyr_down.cu).  The 25 loads of an
// output hit L1 (neighbouring outputs share 15 of them); a shared-memory tiled version (source tile + shared vertical
// sums, 5.3 loads and 11.4 FMAs per output) measured SLOWER on B200: 5.4 us against 3.6 us per launch (round 2).
__global__ void pyrdown_kernel(const float *__restrict__ src, int sw, int sh, float *__restrict__ dst, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    int rows[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) rows[i] = reflect101(2 * y - 2 + i, sh) * sw;
    float col[5];
#pragma unroll
    for (int k = 0; k < 5; ++k)
    {
        const int c = reflect101(2 * x - 2 + k, sw);
        float s = __fmul_rn(0.0625f, __ldg(src + rows[0] + c));
        s = __fmaf_rn(0.25f, __ldg(src + rows[1] + c), s);
        s = __fmaf_rn(0.375f, __ldg(src + rows[2] + c), s);
        s = __fmaf_rn(0.25f, __ldg(src + rows[3] + c), s);
        s = __fmaf_rn(0.0625f, __ldg(src + rows[4] + c), s);
        col[k] = s;
    }
    float s = __fmul_rn(0.0625f, col[0]);
    s = __fmaf_rn(0.25f, col[1], s);
    s = __fmaf_rn(0.375f, col[2], s);
    s = __fmaf_rn(0.25f, col[3], s);
    s = __fmaf_rn(0.0625f, col[4], s);
    dst[y * dw + x] = s;
}

struct FrontArgs
{
    int levels;
    Intr k[KFB_MAX_LEVELS];
    const float *raw[KFB_MAX_LEVELS];
    float *depth[KFB_MAX_LEVELS];
    float4 *v[KFB_MAX_LEVELS];
    float4 *n[KFB_MAX_LEVELS];
    int radius;
    float ss, sc; // -0.5/sigma^2
    float max_dist;
};

#define FT_W 32
#define FT_H 8

// ---- bilateral + scale + truncate -------------------------------------------------------------
__global__ void __launch_bounds__(FT_W *FT_H) bilateral_kernel(const FrontArgs a)
{
    const int l = blockIdx.z;
    const Intr k = a.k[l];
    const int bx = blockIdx.x * FT_W, by = blockIdx.y * FT_H;
    if (bx >= k.w || by >= k.h) return;
    constexpr int R = 2; // the tile is sized for the default 5x5 window; other radii take the global path
    __shared__ float tile[FT_H + 2 * R][FT_W + 2 * R + 1];
    const float *src = a.raw[l];
    const int tid = threadIdx.y * FT_W + threadIdx.x;
    const bool use_tile = (a.radius == R);
    if (use_tile)
    {
        for (int i = tid; i < (FT_H + 2 * R) * (FT_W + 2 * R); i += FT_W * FT_H)
        {
            const int ty = i / (FT_W + 2 * R), tx = i % (FT_W + 2 * R);
            const int gy = reflect101(by + ty - R, k.h), gx = reflect101(bx + tx - R, k.w);
            tile[ty][tx] = __ldg(src + gy * k.w + gx);
        }
        __syncthreads();
    }
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= k.w || y >= k.h) return;
    const int r = a.radius;
    const float r2 = (float)(r * r);
    const float center = use_tile ? tile[threadIdx.y + R][threadIdx.x + R] : __ldg(src + y * k.w + x);
    float sum1 = 0.f, sum2 = 0.f;
    for (int dy = -r; dy <= r; ++dy)
        for (int dx = -r; dx <= r; ++dx)
        {
            const float space2 = (float)(dx * dx + dy * dy);
            if (space2 > r2) continue;
            const float value = use_tile ? tile[threadIdx.y + R + dy][threadIdx.x + R + dx]
                                         : __ldg(src + reflect101(y + dy, k.h) * k.w + reflect101(x + dx, k.w));
            const float ad = fabsf(__fsub_rn(value, center));
            const float wgt = expf(__fmaf_rn(space2, a.ss, __fmul_rn(__fmul_rn(ad, ad), a.sc)));
            sum1 = __fmaf_rn(wgt, value, sum1);
            sum2 = __fadd_rn(sum2, wgt);
        }
    float d = __fmul_rn(__fdiv_rn(sum1, sum2), 0.001f); // depthTruncation: mm -> m
    if (d > a.max_dist) d = 0.f;
    a.depth[l][y * k.w + x] = d;
}

// ---- vertex + normal ------------------------------------------------------------------------------
__device__ __forceinline__ float3 reproj(const Intr &k, float rfx, float rfy, int u, int v, float z)
{
    // device_utils.cuh:22-27: __fdividef(z*(u-cx), fx) = MUFU.RCP(fx) * (z*(u-cx)); NaN depth -> 0
    if (isnan(z)) return make_float3(0.f, 0.f, 0.f);
    return make_float3(__fmul_rn(rfx, __fmul_rn(z, __fsub_rn((float)u, k.cx))),
                       __fmul_rn(rfy, __fmul_rn(z, __fsub_rn((float)v, k.cy))), z);
}

__global__ void __launch_bounds__(FT_W *FT_H) vertex_normal_kernel(const FrontArgs a)
{
    const int l = blockIdx.z;
    const Intr k = a.k[l];
    const int bx = blockIdx.x * FT_W, by = blockIdx.y * FT_H;
    if (bx >= k.w || by >= k.h) return;
    __shared__ float tile[FT_H + 2][FT_W + 2 + 1];
    const float *src = a.depth[l];
    const int tid = threadIdx.y * FT_W + threadIdx.x;
    for (int i = tid; i < (FT_H + 2) * (FT_W + 2); i += FT_W * FT_H)
    {
        const int ty = i / (FT_W + 2), tx = i % (FT_W + 2);
        const int gy = by + ty - 1, gx = bx + tx - 1;
        tile[ty][tx] = (gx >= 0 && gx < k.w && gy >= 0 && gy < k.h) ? src[gy * k.w + gx] : 0.f;
    }
    __syncthreads();
    const int x = bx + threadIdx.x, y = by + threadIdx.y;
    if (x >= k.w || y >= k.h) return;
    const float rfx = rcp_fdividef(k.fx), rfy = rcp_fdividef(k.fy);
    const int tx = threadIdx.x + 1, ty = threadIdx.y + 1;
    const float3 vc = reproj(k, rfx, rfy, x, y, tile[ty][tx]);
    a.v[l][y * k.w + x] = make_float4(vc.x, vc.y, vc.z, 0.f);
    float3 nrm = make_float3(0.f, 0.f, 0.f);
    if (!(x < 1 || x >= k.w - 1 || y < 1 || y >= k.h - 1))
    {
        const float3 vl = reproj(k, rfx, rfy, x - 1, y, tile[ty][tx - 1]);
        const float3 vr = reproj(k, rfx, rfy, x + 1, y, tile[ty][tx + 1]);
        const float3 vu = reproj(k, rfx, rfy, x, y - 1, tile[ty - 1][tx]);
        const float3 vd = reproj(k, rfx, rfy, x, y + 1, tile[ty + 1][tx]);
        float nx = 0.f, ny = 0.f, nz = 0.f;
        if (!(vl.z == 0.f || vr.z == 0.f || vu.z == 0.f || vd.z == 0.f))
        {
            const float ax = __fsub_rn(vl.x, vr.x), ay = __fsub_rn(vl.y, vr.y), az = __fsub_rn(vl.z, vr.z);
            const float bx_ = __fsub_rn(vu.x, vd.x), by_ = __fsub_rn(vu.y, vd.y), bz = __fsub_rn(vu.z, vd.z);
            nx = __fmaf_rn(ay, bz, -__fmul_rn(az, by_));
            ny = __fmaf_rn(az, bx_, -__fmul_rn(ax, bz));
            nz = __fmaf_rn(ax, by_, -__fmul_rn(ay, bx_));
            if (nz > 0.f) { nx = -nx; ny = -ny; nz = -nz; }
        }
        // normalize(): IEEE sqrt and divisions; 0/0 = NaN marks the invalid normal (device_types.hpp:253-257)
        const float t = __fsqrt_rn(dot3c(nx, ny, nz, nx, ny, nz));
        nrm = make_float3(__fdiv_rn(nx, t), __fdiv_rn(ny, t), __fdiv_rn(nz, t));
    }
    a.n[l][y * k.w + x] = make_float4(nrm.x, nrm.y, nrm.z, 0.f);
}

// ---- model-map pyramid (resizePointsNormals) ----------------------------------------------------------
__global__ void resize_maps_kernel(const float4 *__restrict__ vb, const float4 *__restrict__ nb, int bw,
                                   float4 *__restrict__ vs, float4 *__restrict__ ns, int sw, int sh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= sw || y >= sh) return;
    const int i00 = (2 * y) * bw + 2 * x;
    const float4 d00 = vb[i00], d01 = vb[i00 + 1], d10 = vb[i00 + bw], d11 = vb[i00 + bw + 1];
    float4 vo = make_float4(0.f, 0.f, 0.f, 0.f), no = vo;
    if (!isnan(__fmul_rn(__fmul_rn(__fmul_rn(d00.x, d01.x), d10.x), d11.x)))
    {
        vo.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(d00.x, d01.x), d10.x), d11.x), 0.25f);
        vo.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(d00.y, d01.y), d10.y), d11.y), 0.25f);
        vo.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(d00.z, d01.z), d10.z), d11.z), 0.25f);
        const float4 n00 = nb[i00], n01 = nb[i00 + 1], n10 = nb[i00 + bw], n11 = nb[i00 + bw + 1];
        no.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n00.x, n01.x), n10.x), n11.x), 0.25f);
        no.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n00.y, n01.y), n10.y), n11.y), 0.25f);
        no.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n00.z, n01.z), n10.z), n11.z), 0.25f);
    }
    vs[y * sw + x] = vo;
    ns[y * sw + x] = no;
}

// ---- float3 <-> float4 converters for the download/upload hooks ------------------------------------------
__global__ void map4to3_kernel(const float4 *__restrict__ src, float *__restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = src[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}
__global__ void map3to4_kernel(const float *__restrict__ src, float4 *__restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}
int launch_map_convert(kfb_ctx *ctx, const float4 *src, float *dst3, size_t n)
{
    map4to3_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src, dst3, n);
    KFB_LAUNCH_CHECK(ctx);
    return KFB_OK;
}
int launch_map_convert_in(kfb_ctx *ctx, const float *src3, float4 *dst, size_t n)
{
    map3to4_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(src3, dst, n);
    KFB_LAUNCH_CHECK(ctx);
    return KFB_OK;
}

int launch_frontend(kfb_ctx *ctx)
{
    int rcf = fork_front(ctx);
    if (rcf) return rcf;
    const int L = ctx->levels;
    for (int l = 1; l < L; ++l)
    {
        const Intr &s = ctx->L[l - 1].k, &d = ctx->L[l].k;
        dim3 b(32, 8), g((d.w + 31) / 32, (d.h + 7) / 8);
        pyrdown_kernel<<<g, b, 0, ctx->fstream>>>(ctx->L[l - 1].raw, s.w, s.h, ctx->L[l].raw, d.w, d.h);
        KFB_LAUNCH_CHECK(ctx);
    }
    FrontArgs a;
    a.levels = L;
    for (int l = 0; l < L; ++l)
    {
        a.k[l] = ctx->L[l].k;
        a.raw[l] = ctx->L[l].raw;
        a.depth[l] = ctx->L[l].depth;
        a.v[l] = ctx->L[l].v[ctx->cur];
        a.n[l] = ctx->L[l].n[ctx->cur];
    }
    // host-side parameter preparation of cv::cuda::bilateralFilter
    float sc = ctx->p.bfilter_color_sigma, ss = ctx->p.bfilter_spatial_sigma;
    if (sc <= 0) sc = 1;
    if (ss <= 0) ss = 1;
    int radius = ctx->p.bfilter_kernel_size <= 0 ? (int)lrintf(ss * 1.5f) : ctx->p.bfilter_kernel_size / 2;
    if (radius < 1) radius = 1;
    a.radius = radius;
    a.ss = -0.5f / (ss * ss);
    a.sc = -0.5f / (sc * sc);
    a.max_dist = ctx->p.dfilter_dist;
    const Intr &k0 = ctx->L[0].k;
    dim3 b(FT_W, FT_H), g((k0.w + FT_W - 1) / FT_W, (k0.h + FT_H - 1) / FT_H, L);
    bilateral_kernel<<<g, b, 0, ctx->fstream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    vertex_normal_kernel<<<g, b, 0, ctx->fstream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    // integrate's per-pixel tables of this depth image, once the previous frame's integrate has read the old ones
    KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->fstream, ctx->ev_tables_free, 0));
    const int rct = launch_build_tables(ctx, ctx->fstream);
    if (rct) return rct;
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_front, ctx->fstream));
    ctx->front_pending = 1;
    return KFB_OK;
}

__global__ void u16_to_f32_kernel(const uint16_t *__restrict__ src, float *__restrict__ dst, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (float)src[i]; // exact: the sensor's millimetres, as depth_sensor.cpp:192 converts them
}
int launch_u16_to_f32(kfb_ctx *ctx, const uint16_t *src, float *dst, size_t n, cudaStream_t stream)
{
    u16_to_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, dst, n);
    KFB_LAUNCH_CHECK(ctx);
    return KFB_OK;
}

int launch_model_pyramid(kfb_ctx *ctx)
{
    if (ctx->pyramid_fresh) { ctx->pyramid_fresh = 0; return KFB_OK; } // done by the raycast epilogue
    for (int l = 1; l < ctx->levels; ++l)
    {
        const Intr &s = ctx->L[l - 1].k, &d = ctx->L[l].k;
        dim3 b(32, 8), g((d.w + 31) / 32, (d.h + 7) / 8);
        resize_maps_kernel<<<g, b, 0, ctx->stream>>>(ctx->L[l - 1].v[ctx->prev], ctx->L[l - 1].n[ctx->prev], s.w,
                                                    ctx->L[l].v[ctx->prev], ctx->L[l].n[ctx->prev], d.w, d.h);
        KFB_LAUNCH_CHECK(ctx);
    }
    return KFB_OK;
}

} // namespace kfb
