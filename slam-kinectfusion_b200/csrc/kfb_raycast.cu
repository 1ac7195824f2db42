// kfb_raycast.cu -- TSDF raycast (replaces kf::device::raycast, kfusion/src/tsdf_volume.cu:120-273)
// and the render kernels (kf::device::renderPhong / renderNormals, image_process.cu:137-221).
//
// Semantics are the reference's, rounding for rounding: one-voxel steps,
// nearest-neighbour sample for the sign test, first +->- pair => vertex at
// ray_len -/+ vs*cur/(cur-next) (the reference's sign quirk behind compat_raycast_ts_sign),
// normal = normalised central difference of six trilinear fetches, outputs rotated back to the
// camera frame, explicit zeros on a miss.  Structure is new: a warp owns an 8x4 pixel tile (coherent
// rays), the march is software-pipelined in batches (sample positions do not depend on fetched
// values, so KFB_RC_BATCH independent gathers are in flight per ray before the first sign test),
// and the running sums `nextp += dir*voxel_size`, `ray_len += step` are replayed exactly.
#include "kfb_common.cuh"

namespace kfb
{

struct RaycastArgs
{
    const uint32_t *vol;
    int X, Y, Z;          // global dims
    int z_store0;         // first stored plane
    Pose pose;            // cam2vol
    Mat3 rinv;
    Intr k;
    float vs[3], vsinv[3], gd[3], range[3];
    float step_len;
    float4 *vmap, *nmap;
    float *hit_t;
    int ts_sign_compat;
};

#define KFB_RC_BATCH 4

__device__ __forceinline__ float vox_tsdf(const RaycastArgs &a, int x, int y, int z)
{
    const size_t i = ((size_t)(z - a.z_store0) * a.Y + y) * a.X + x;
    const short s = __ldg(reinterpret_cast<const short *>(a.vol + i)); // low half = tsdf
    return __fmul_rn((float)s, KFB_DIVSHORTMAX);
}
// raycasthelper::voxel2tsdf (tsdf_volume.cu:178-191)
__device__ __forceinline__ float fetch_nn(const RaycastArgs &a, float px, float py, float pz)
{
    const int x = __float2int_rn(__fmul_rn(px, a.vsinv[0]));
    const int y = __float2int_rn(__fmul_rn(py, a.vsinv[1]));
    const int z = __float2int_rn(__fmul_rn(pz, a.vsinv[2]));
    if (x >= a.X - 1 || y >= a.Y - 1 || z >= a.Z - 1 || x < 1 || y < 1 || z < 1) return __int_as_float(0x7fffffff);
    return vox_tsdf(a, x, y, z);
}
// interpolate (tsdf_volume.cu:137-161)
__device__ float interp(const RaycastArgs &a, float fx, float fy, float fz)
{
    const int gx = __float2int_rd(fx), gy = __float2int_rd(fy), gz = __float2int_rd(fz);
    if (gx < 0 || gx >= a.X - 1 || gy < 0 || gy >= a.Y - 1 || gz < 0 || gz >= a.Z - 1) return __int_as_float(0x7fffffff);
    const float fa = __fsub_rn(fx, (float)gx), fb = __fsub_rn(fy, (float)gy), fc = __fsub_rn(fz, (float)gz);
    const float a1 = __fsub_rn(1.f, fa), b1 = __fsub_rn(1.f, fb), c1 = __fsub_rn(1.f, fc);
    const float v000 = vox_tsdf(a, gx, gy, gz), v001 = vox_tsdf(a, gx, gy, gz + 1);
    const float v010 = vox_tsdf(a, gx, gy + 1, gz), v011 = vox_tsdf(a, gx, gy + 1, gz + 1);
    const float v100 = vox_tsdf(a, gx + 1, gy, gz), v101 = vox_tsdf(a, gx + 1, gy, gz + 1);
    const float v110 = vox_tsdf(a, gx + 1, gy + 1, gz), v111 = vox_tsdf(a, gx + 1, gy + 1, gz + 1);
    float t = 0.f;
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v000, a1), b1), c1, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v001, a1), b1), fc, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v010, a1), fb), c1, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v011, a1), fb), fc, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v100, fa), b1), c1, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v101, fa), b1), fc, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v110, fa), fb), c1, t);
    t = __fmaf_rn(__fmul_rn(__fmul_rn(v111, fa), fb), fc, t);
    return t;
}

__global__ void __launch_bounds__(128) raycast_kernel(const RaycastArgs a)
{
    // warp = 8x4 pixel tile; block = 8x16 pixels
    const int x = blockIdx.x * 8 + threadIdx.x;
    const int y = blockIdx.y * 16 + threadIdx.y;
    if (x >= a.k.w || y >= a.k.h) return;
    const int pix = y * a.k.w + x;
    float4 vout = make_float4(0.f, 0.f, 0.f, 0.f), nout = vout;
    float t_hit = __int_as_float(0x7f800000); // +inf = miss

    const float ox = a.pose.t[0], oy = a.pose.t[1], oz = a.pose.t[2];
    // reproj(x, y, 1) and ray direction (tsdf_volume.cu:217-220)
    const float px = __fmul_rn(rcp_fdividef(a.k.fx), __fsub_rn((float)x, a.k.cx));
    const float py = __fmul_rn(rcp_fdividef(a.k.fy), __fsub_rn((float)y, a.k.cy));
    float dx = __fadd_rn(__fmaf_rn(px, a.pose.R.m[0], __fmul_rn(py, a.pose.R.m[1])), a.pose.R.m[2]);
    float dy = __fadd_rn(__fmaf_rn(px, a.pose.R.m[3], __fmul_rn(py, a.pose.R.m[4])), a.pose.R.m[5]);
    float dz = __fadd_rn(__fmaf_rn(px, a.pose.R.m[6], __fmul_rn(py, a.pose.R.m[7])), a.pose.R.m[8]);
    {
        const float t = __fsqrt_rn(dot3c(dx, dy, dz, dx, dy, dz));
        dx = __fdividef(dx, t); dy = __fdividef(dy, t); dz = __fdividef(dz, t);
    }
    // intersect() (tsdf_volume.cu:120-136)
    const float ix = __fdiv_rn(1.f, dx), iy = __fdiv_rn(1.f, dy), iz = __fdiv_rn(1.f, dz);
    const float bx = __fmul_rn(ix, __fsub_rn(0.f, ox)), by = __fmul_rn(iy, __fsub_rn(0.f, oy)), bz = __fmul_rn(iz, __fsub_rn(0.f, oz));
    const float tx = __fmul_rn(ix, __fsub_rn(a.range[0], ox)), ty = __fmul_rn(iy, __fsub_rn(a.range[1], oy)),
                tz = __fmul_rn(iz, __fsub_rn(a.range[2], oz));
    const float mnx = fminf(tx, bx), mny = fminf(ty, by), mnz = fminf(tz, bz);
    const float mxx = fmaxf(tx, bx), mxy = fmaxf(ty, by), mxz = fmaxf(tz, bz);
    const float tnear = fmaxf(fmaxf(mnx, mny), fmaxf(mnx, mnz));
    const float tfar = fminf(fminf(mxx, mxy), fminf(mxx, mxz));
    float ray_len = fmaxf(tnear, 0.f);
    if (!(ray_len >= tfar))
    {
        ray_len = __fadd_rn(ray_len, a.step_len);
        float nx = __fmaf_rn(dx, ray_len, ox), ny = __fmaf_rn(dy, ray_len, oy), nz = __fmaf_rn(dz, ray_len, oz);
        float tnext = fetch_nn(a, nx, ny, nz);
        bool done = false;
        while (!done && ray_len < tfar)
        {
            float val[KFB_RC_BATCH];
#pragma unroll
            for (int b = 0; b < KFB_RC_BATCH; ++b)
            {
                nx = __fmaf_rn(dx, a.vs[0], nx);
                ny = __fmaf_rn(dy, a.vs[1], ny);
                nz = __fmaf_rn(dz, a.vs[2], nz);
                val[b] = fetch_nn(a, nx, ny, nz);
            }
#pragma unroll
            for (int b = 0; b < KFB_RC_BATCH; ++b)
            {
                if (!done && ray_len < tfar)
                {
                    const float tcur = tnext;
                    tnext = val[b];
                    if (!isnan(tnext))
                    {
                        if (tcur < 0.f && tnext > 0.f)
                            done = true; // back face: stop, no hit (tsdf_volume.cu:242-243)
                        else if (tcur > 0.f && tnext < 0.f)
                        {
                            const float q = rcp_fdividef(__fsub_rn(tcur, tnext));
                            const float num = __fmul_rn(tcur, a.vs[0]);
                            const float Ts = a.ts_sign_compat ? __fmaf_rn(q, -num, ray_len) : __fmaf_rn(q, num, ray_len);
                            const float vx = __fmaf_rn(dx, Ts, ox), vy = __fmaf_rn(dy, Ts, oy), vz = __fmaf_rn(dz, Ts, oz);
                            // compute_normal (tsdf_volume.cu:192-209)
                            const float sx = __fmul_rn(vx, a.vsinv[0]), sy = __fmul_rn(vy, a.vsinv[1]), sz = __fmul_rn(vz, a.vsinv[2]);
                            const float Fx1 = interp(a, __fmul_rn(__fadd_rn(vx, a.gd[0]), a.vsinv[0]), sy, sz);
                            const float Fx2 = interp(a, __fmul_rn(__fsub_rn(vx, a.gd[0]), a.vsinv[0]), sy, sz);
                            const float Fy1 = interp(a, sx, __fmul_rn(__fadd_rn(vy, a.gd[1]), a.vsinv[1]), sz);
                            const float Fy2 = interp(a, sx, __fmul_rn(__fsub_rn(vy, a.gd[1]), a.vsinv[1]), sz);
                            const float Fz1 = interp(a, sx, sy, __fmul_rn(__fadd_rn(vz, a.gd[2]), a.vsinv[2]));
                            const float Fz2 = interp(a, sx, sy, __fmul_rn(__fsub_rn(vz, a.gd[2]), a.vsinv[2]));
                            float gx = __fdividef(__fsub_rn(Fx1, Fx2), a.gd[0]);
                            float gy = __fdividef(__fsub_rn(Fy1, Fy2), a.gd[1]);
                            float gz = __fdividef(__fsub_rn(Fz1, Fz2), a.gd[2]);
                            const float t = __fsqrt_rn(dot3c(gx, gy, gz, gx, gy, gz));
                            gx = __fdividef(gx, t); gy = __fdividef(gy, t); gz = __fdividef(gz, t);
                            if (!isnan(__fmul_rn(__fmul_rn(gx, gy), gz)))
                            {
                                const float3 nn = rot3(a.rinv, gx, gy, gz);
                                const float3 vv = rot3(a.rinv, __fsub_rn(vx, ox), __fsub_rn(vy, oy), __fsub_rn(vz, oz));
                                nout = make_float4(nn.x, nn.y, nn.z, 0.f);
                                vout = make_float4(vv.x, vv.y, vv.z, 0.f);
                                t_hit = Ts;
                                done = true;
                            }
                        }
                    }
                    ray_len = __fadd_rn(ray_len, a.step_len);
                }
            }
        }
    }
    a.vmap[pix] = vout;
    a.nmap[pix] = nout;
    if (a.hit_t) a.hit_t[pix] = t_hit;
}

int launch_raycast(kfb_ctx *ctx, const float cam2vol12[12], const float rinv9[9])
{
    RaycastArgs a;
    a.vol = ctx->vol;
    a.X = ctx->p.volu_dims[0]; a.Y = ctx->p.volu_dims[1]; a.Z = ctx->p.volu_dims[2];
    a.z_store0 = ctx->z0;
    a.pose = make_pose(cam2vol12);
    for (int i = 0; i < 9; ++i) a.rinv.m[i] = rinv9[i];
    a.k = ctx->L[0].k;
    for (int i = 0; i < 3; ++i)
    {
        a.vs[i] = ctx->voxel_size[i];
        a.vsinv[i] = 1.f / ctx->voxel_size[i];  // raycasthelper ctor, host IEEE division (tsdf_volume.cu:176)
        a.gd[i] = ctx->voxel_size[i] * 0.5f;    // tsdf_volume.cu:175
        a.range[i] = ctx->p.volu_range[i];
    }
    a.step_len = ctx->voxel_size[0];            // tsdf_volume.cu:174
    a.vmap = ctx->L[0].v[ctx->prev];
    a.nmap = ctx->L[0].n[ctx->prev];
    a.hit_t = ctx->hit_t;
    a.ts_sign_compat = ctx->p.compat_raycast_ts_sign;
    dim3 block(8, 16), grid((a.k.w + 7) / 8, (a.k.h + 15) / 16);
    if (ctx->profiling) cudaEventRecord(ctx->events[58], ctx->stream);
    raycast_kernel<<<grid, block, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    if (ctx->profiling) cudaEventRecord(ctx->events[59], ctx->stream);
    return KFB_OK;
}

// ---- render (image_process.cu:137-221) ------------------------------------------------------------
__global__ void render_normals_kernel(const float4 *__restrict__ nmap, uint8_t *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = nmap[i];
    out[3 * i + 0] = (uint8_t)(int)__fmul_rn(fabsf(v.x), 255.f);
    out[3 * i + 1] = (uint8_t)(int)__fmul_rn(fabsf(v.y), 255.f);
    out[3 * i + 2] = (uint8_t)(int)__fmul_rn(fabsf(v.z), 255.f);
}

__global__ void render_phong_kernel(const float4 *__restrict__ vmap, const float4 *__restrict__ nmap, float ex,
                                    float ey, float ez, uint8_t *__restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = vmap[i], nn = nmap[i];
    if (nn.x == 0 && nn.y == 0 && nn.z == 0) return;
    if (v.x == 0 && v.y == 0 && v.z == 0) return;
    float3 e = make_float3(__fsub_rn(ex, v.x), __fsub_rn(ey, v.y), __fsub_rn(ez, v.z));
    float3 l = make_float3(__fsub_rn(500.f, v.x), __fsub_rn(500.f, v.y), __fsub_rn(-500.f, v.z));
    float t = __fsqrt_rn(dot3c(e.x, e.y, e.z, e.x, e.y, e.z));
    e = make_float3(__fdividef(e.x, t), __fdividef(e.y, t), __fdividef(e.z, t));
    t = __fsqrt_rn(dot3c(l.x, l.y, l.z, l.x, l.y, l.z));
    l = make_float3(__fdividef(l.x, t), __fdividef(l.y, t), __fdividef(l.z, t));
    float lc = dot3c(nn.x, nn.y, nn.z, l.x, l.y, l.z);
    if (lc <= 0) lc = -lc;
    const float li = (float)0.9;
    float coef = __fmul_rn(li, lc);
    const float dfx = __fmul_rn(0.3843f, coef), dfy = __fmul_rn(0.4745f, coef), dfz = __fmul_rn(0.580f, coef);
    float3 hh = make_float3(__fadd_rn(l.x, e.x), __fadd_rn(l.y, e.y), __fadd_rn(l.z, e.z));
    t = __fsqrt_rn(dot3c(hh.x, hh.y, hh.z, hh.x, hh.y, hh.z));
    hh = make_float3(__fdividef(hh.x, t), __fdividef(hh.y, t), __fdividef(hh.z, t));
    float hc = dot3c(nn.x, nn.y, nn.z, hh.x, hh.y, hh.z);
    if (hc < 0) hc = -hc;
    coef = __fmul_rn(li, powf(hc, 10.f));
    const float sp = (float)(0.5 * (double)coef);
    const float kx = fminf(1.f, __fadd_rn(__fadd_rn(0.1f, dfx), sp));
    const float ky = fminf(1.f, __fadd_rn(__fadd_rn(0.1f, dfy), sp));
    const float kz = fminf(1.f, __fadd_rn(__fadd_rn(0.1f, dfz), sp));
    out[3 * i + 0] = (uint8_t)(int)__fmul_rn(kx, 255.f);
    out[3 * i + 1] = (uint8_t)(int)__fmul_rn(ky, 255.f);
    out[3 * i + 2] = (uint8_t)(int)__fmul_rn(kz, 255.f);
}

int launch_render(kfb_ctx *ctx, int phong, const float eye3[3], uint8_t *host_bgr)
{
    const Intr &k = ctx->L[0].k;
    const int n = k.w * k.h;
    // the reference renders into pframe->cmap, zeroed by pframe->reset() each frame (kinectfusion.cpp:112)
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->render_dev, 0, (size_t)n * 3, ctx->stream));
    if (phong)
        render_phong_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->L[0].v[ctx->prev], ctx->L[0].n[ctx->prev],
                                                                   eye3[0], eye3[1], eye3[2], ctx->render_dev, n);
    else
        render_normals_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->L[0].n[ctx->prev], ctx->render_dev, n);
    KFB_LAUNCH_CHECK(ctx);
    KFB_CUDA(ctx, cudaMemcpyAsync(ctx->render_host, ctx->render_dev, (size_t)n * 3, cudaMemcpyDeviceToHost, ctx->stream));
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(host_bgr, ctx->render_host, (size_t)n * 3);
    return KFB_OK;
}

} // namespace kfb
