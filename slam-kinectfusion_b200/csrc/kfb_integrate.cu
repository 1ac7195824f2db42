// kfb_integrate.cu -- TSDF integration (replaces kf::device::integrate,
// kfusion/src/tsdf_volume.cu:34-111) and volume reset (tsdf_volume.cu:11-32).
//
// Voxel-column sweep over the packed {int16 tsdf, int16 weight} volume: a thread owns
// four consecutive x voxels and marches z; one 128-bit load and one 128-bit store per
// z step, issued only when at least one of the four voxels passes the reference's
// update predicate (so culled voxels cost no HBM traffic).  The predicate and the
// update arithmetic are the reference's, rounding for rounding:
//   * vc(z) is the reference's running sum, vc = fma(voxel_size.x, R[:,2], vc), replayed
//     from z = 1 (also across z-chunks / z-slabs) so every plane sees identical bits;
//   * pixel = round-half-even(fma(vc.x * MUFU.RCP(vc.z), fx, cx)), done with the 2^23
//     magic-number add so it stays off the conversion pipe;
//   * sdf = depth - |vc| / lambda is only evaluated exactly inside a narrow band around
//     the surface: a per-pixel table of conservative |vc|^2 thresholds (fp16, rounded
//     outward) classifies "certainly free space => tsdf == 1.0f exactly" and "certainly
//     behind the surface => rejected" without the two sqrt/rcp chains.  The thresholds
//     are proved conservative in build_tables_kernel, so classification never changes a
//     result, only skips work.
#include "kfb_common.cuh"

namespace kfb
{

struct IntegrateArgs
{
    uint32_t *vol;
    int X, Y;
    int z_store0;   // global z of the first stored plane
    int zb, ze;     // global planes to process: [zb, ze), zb >= 1
    int zchunk;     // planes per blockIdx.z
    Pose pose;      // vol2cam
    float vsx, vsy, vsz;
    float trunc;
    float fx, fy, cx, cy;
    int w, h;
    const __half2 *thr;
    const float2 *exact;
    int max_weight;
    unsigned long long *counter;
};

#define KFB_MAGIC_F 12582912.0f   // 1.5 * 2^23
#define KFB_MAGIC_I 0x4B400000
#define KFB_SKIP (-4.0f)

// ---- per-pixel tables ---------------------------------------------------------------
// exact[p] = {depth, MUFU.RCP(lambda)} with lambda = sqrt(((u-cx)/fx)^2 + ((v-cy)/fy)^2 + 1)
// computed with the reference's operations (tsdf_volume.cu:65-68, device_utils.cuh:22-27).
// thr[p]   = {hi2, lo2}:  |vc|^2 <= hi2  ==> the reference's tsdf is exactly 1.0f
//                         |vc|^2 >  lo2  ==> the reference rejects the voxel (sdf < -trunc)
// Proof sketch (all quantities positive, eps = 2^-24):
//   nrm = sqrt_rn(d2) in sqrt(d2)(1 +- eps); il = MUFU.RCP(lambda) in (1/lambda)(1 +- 2^-22);
//   nsdf = RN(il*nrm - depth).  With T = trunc(1+1e-5):
//   d2 <= ((depth - T - 1e-6) lambda)^2 (1-8e-6) => il*nrm <= depth - T (reals) => -nsdf >= T
//        => MUFU.RCP(trunc) * (-nsdf) >= (1+1e-5)(1-2^-22)(1-eps) > 1 => fmin(1, .) == 1.
//   d2 >  ((depth + T + 1e-6) lambda)^2 (1+8e-6) => il*nrm - depth > T > trunc => nsdf > trunc.
//   The 1e-6 absorbs the float rounding of depth -+ T (depth < 8 m); fp16 conversion rounds
//   hi2 down and lo2 up.  Invalid depth (<= 0 or NaN) => {-1, -1}: everything rejected, as the
//   reference does (`depth <= 0` skip; NaN depth makes sdf NaN, which fails `sdf >= -trunc`).
__global__ void build_tables_kernel(const float *__restrict__ depth, int w, int h, float fx, float fy, float cx,
                                    float cy, float trunc, __half2 *__restrict__ thr, float2 *__restrict__ exact)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= w || v >= h) return;
    const int p = v * w + u;
    const float d = depth[p];
    const float lx = __fmul_rn(rcp_fdividef(fx), __fsub_rn((float)u, cx));
    const float ly = __fmul_rn(rcp_fdividef(fy), __fsub_rn((float)v, cy));
    const float lam = __fsqrt_rn(__fadd_rn(__fmaf_rn(lx, lx, __fmul_rn(ly, ly)), 1.0f));
    const float il = rcp_fdividef(lam);
    exact[p] = make_float2(d, il);
    float hi2 = -1.f, lo2 = -1.f;
    if (d > 0.f)
    {
        const float T = trunc * 1.00001f;
        const float a = d - T - 1e-6f;
        if (a > 0.f)
        {
            const float b = a * lam;
            hi2 = b * b * (1.f - 8e-6f);
        }
        const float c = (d + T + 1e-6f) * lam;
        lo2 = c * c * (1.f + 8e-6f);
    }
    thr[p] = __halves2half2(__float2half_rd(hi2), __float2half_ru(lo2));
}

// ---- the sweep ------------------------------------------------------------------------
template <int U, bool COUNT>
__global__ void __launch_bounds__(128) integrate_kernel(const IntegrateArgs a)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x0 >= a.X || y >= a.Y) return;

    const int zstart = a.zb + blockIdx.z * a.zchunk;
    const int zend = min(zstart + a.zchunk, a.ze);
    if (zstart >= zend) return;

    // vc at z = 0: R * (x*vs.x, y*vs.y, 0*vs.z) + t   (tsdf_volume.cu:49-50)
    float vx[4], vy[4], vz[4];
    {
        const float py = __fmul_rn((float)y, a.vsy);
        const float pz = __fmul_rn(0.f, a.vsz);
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const float px = __fmul_rn((float)(x0 + k), a.vsx);
            const float3 r = rot3(a.pose.R, px, py, pz);
            vx[k] = __fadd_rn(r.x, a.pose.t[0]);
            vy[k] = __fadd_rn(r.y, a.pose.t[1]);
            vz[k] = __fadd_rn(r.z, a.pose.t[2]);
        }
    }
    const float sx = a.pose.R.m[2], sy = a.pose.R.m[5], sz = a.pose.R.m[8];
    // replay of the reference's running sum up to the first plane of this chunk (tsdf_volume.cu:56)
    for (int z = 1; z < zstart; ++z)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            vx[k] = __fmaf_rn(a.vsx, sx, vx[k]);
            vy[k] = __fmaf_rn(a.vsx, sy, vy[k]);
            vz[k] = __fmaf_rn(a.vsx, sz, vz[k]);
        }
    }

    const float rtrunc = rcp_fdividef(a.trunc);
    const size_t plane4 = ((size_t)a.X * a.Y) >> 2; // uint4 per plane
    uint4 *vp = reinterpret_cast<uint4 *>(a.vol) + ((size_t)(zstart - a.z_store0) * a.Y + y) * (a.X >> 2) + (x0 >> 2);
    unsigned int n_upd = 0;

    for (int z = zstart; z < zend; z += U)
    {
        float ts[U][4];
        uint4 word[U];
        bool need[U];
        // ---- phase A: advance, classify, issue loads -------------------------------------
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            need[u] = false;
            if (z + u < zend)
            {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    vx[k] = __fmaf_rn(a.vsx, sx, vx[k]);
                    vy[k] = __fmaf_rn(a.vsx, sy, vy[k]);
                    vz[k] = __fmaf_rn(a.vsx, sz, vz[k]);
                    float t = KFB_SKIP;
                    const float cz = vz[k];
                    if (!(cz <= 0.f))
                    {
                        float qx, qy;
                        if (cz >= KFB_FLT_MIN)
                        {
                            const float r = mufu_rcp(cz);
                            qx = __fmul_rn(r, vx[k]);
                            qy = __fmul_rn(r, vy[k]);
                        }
                        else
                        {
                            qx = __fdividef(vx[k], cz);
                            qy = __fdividef(vy[k], cz);
                        }
                        const int ui = __float_as_int(__fadd_rn(__fmaf_rn(qx, a.fx, a.cx), KFB_MAGIC_F)) - KFB_MAGIC_I;
                        const int vi = __float_as_int(__fadd_rn(__fmaf_rn(qy, a.fy, a.cy), KFB_MAGIC_F)) - KFB_MAGIC_I;
                        if ((unsigned)ui < (unsigned)a.w && (unsigned)vi < (unsigned)a.h)
                        {
                            const int p = vi * a.w + ui;
                            const float2 th = __half22float2(__ldg(a.thr + p));
                            const float d2 = dot3c(vx[k], vy[k], cz, vx[k], vy[k], cz);
                            if (d2 <= th.x)
                                t = 1.0f;
                            else if (!(d2 > th.y))
                            {
                                const float2 e = __ldg(a.exact + p);
                                const float nsdf = __fmaf_rn(e.y, __fsqrt_rn(d2), -e.x);
                                if (nsdf <= a.trunc) t = fminf(1.f, __fmul_rn(rtrunc, -nsdf));
                            }
                        }
                    }
                    ts[u][k] = t;
                    need[u] = need[u] || (t != KFB_SKIP);
                }
                if (need[u]) word[u] = __ldcs(vp + (size_t)u * plane4);
            }
        }
        // ---- phase B: running weighted mean, re-encode, store (tsdf_volume.cu:69-79) --------
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            if (need[u])
            {
                unsigned int wv[4] = {word[u].x, word[u].y, word[u].z, word[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    if (ts[u][k] != KFB_SKIP)
                    {
                        const int tsv = (int)(short)(wv[k] & 0xffffu);
                        const int wt = (int)(short)(wv[k] >> 16);
                        const float pre = __fmul_rn((float)tsv, KFB_DIVSHORTMAX);
                        const int wp1 = wt + 1;
                        const float rd = rcp_fdividef((float)wp1);
                        const float nt = __fmul_rn(rd, __fmaf_rn(pre, (float)wt, ts[u][k]));
                        int q = __float2int_rz(__fmul_rn(nt, (float)KFB_SHORTMAX));
                        q = max(-KFB_SHORTMAX, min(KFB_SHORTMAX, q));
                        const int nw = min(wp1, a.max_weight);
                        wv[k] = ((unsigned)q & 0xffffu) | ((unsigned)nw << 16);
                        if (COUNT) ++n_upd;
                    }
                }
                __stcs(vp + (size_t)u * plane4, make_uint4(wv[0], wv[1], wv[2], wv[3]));
            }
        }
        vp += (size_t)U * plane4;
    }
    if (COUNT)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_upd += __shfl_xor_sync(0xffffffffu, n_upd, o);
        if ((threadIdx.x & 31) == 0 && n_upd) atomicAdd(a.counter, (unsigned long long)n_upd);
    }
}

int launch_integrate(kfb_ctx *ctx, const float vol2cam12[12], uint64_t *n_updated)
{
    const Intr &k = ctx->L[0].k;
    {
        dim3 b(32, 8), g((k.w + 31) / 32, (k.h + 7) / 8);
        build_tables_kernel<<<g, b, 0, ctx->stream>>>(ctx->L[0].depth, k.w, k.h, k.fx, k.fy, k.cx, k.cy,
                                                     ctx->p.volu_trun_dist, ctx->tab_thr, ctx->tab_exact);
        KFB_LAUNCH_CHECK(ctx);
    }
    IntegrateArgs a;
    a.vol = ctx->vol;
    a.X = ctx->p.volu_dims[0];
    a.Y = ctx->p.volu_dims[1];
    a.z_store0 = ctx->z0;
    a.zb = ctx->z0 < 1 ? 1 : ctx->z0;
    a.ze = ctx->z1;
    a.pose = make_pose(vol2cam12);
    a.vsx = ctx->voxel_size[0]; a.vsy = ctx->voxel_size[1]; a.vsz = ctx->voxel_size[2];
    a.trunc = ctx->p.volu_trun_dist;
    a.fx = k.fx; a.fy = k.fy; a.cx = k.cx; a.cy = k.cy;
    a.w = k.w; a.h = k.h;
    a.thr = ctx->tab_thr;
    a.exact = ctx->tab_exact;
    a.max_weight = ctx->p.tsdf_max_weight;
    a.counter = ctx->counters;

    const int planes = a.ze - a.zb;
    if (planes <= 0) return KFB_OK;
    // z-chunking trades replayed running-sum adds for resident warps; small volumes need it
    // to fill 148 SMs.  KFB_INTEGRATE_ZCHUNKS overrides for tuning.
    const long cols = ((long)(a.X + 127) / 128) * ((a.Y + 3) / 4);
    int zc = 1;
    while (cols * zc < 148L * 8 && zc < 16 && planes / (zc * 2) >= 32) zc *= 2;
    if (const char *e = getenv("KFB_INTEGRATE_ZCHUNKS")) zc = atoi(e) > 0 ? atoi(e) : zc;
    a.zchunk = (planes + zc - 1) / zc;
    dim3 block(32, 4), grid((a.X + 127) / 128, (a.Y + 3) / 4, zc);
    if (n_updated)
    {
        KFB_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, sizeof(unsigned long long), ctx->stream));
        integrate_kernel<2, true><<<grid, block, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        KFB_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, sizeof(unsigned long long),
                                      cudaMemcpyDeviceToHost, ctx->stream));
        KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *n_updated = ctx->counters_host[0];
    }
    else
    {
        if (ctx->profiling) cudaEventRecord(ctx->events[60], ctx->stream);
        integrate_kernel<2, false><<<grid, block, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        if (ctx->profiling) cudaEventRecord(ctx->events[61], ctx->stream);
    }
    return KFB_OK;
}

// resetVolume: zero every stored voxel (the reference's fixed 32x32 grid is a bug)
int launch_reset_volume(kfb_ctx *ctx)
{
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->vol, 0, ctx->vol_voxels * sizeof(uint32_t), ctx->stream));
    return KFB_OK;
}

} // namespace kfb
