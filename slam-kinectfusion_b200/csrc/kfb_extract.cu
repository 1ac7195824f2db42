// kfb_extract.cu -- point-cloud extraction (replaces kf::device::extract_points /
// FullScan6, kfusion/src/tsdf_volume.cu:282-499, and the device side of
// TSDFVolume::fetchPointCloud, kfusion/src/tsdf_volume.cpp:63-84).
//
// Full-volume zero-crossing scan: a thread marches z for one (x, y) column keeping the z+1
// voxel in a register for the next step (each voxel is fetched from HBM once; the +x / +y
// neighbours come from L1/L2), emits up to three points per voxel and appends them with one
// warp-aggregated atomicAdd per step.  Arithmetic follows the reference:
// voxel-centre (+0.5) convention, W != 0 && F != 1 filter, linear interpolation of the
// crossing, transform by volume_pose.  Output order is unspecified (as in the reference);
// compare as sorted sets.
#include "kfb_common.cuh"

namespace kfb
{

struct ExtractArgs
{
    const uint32_t *vol; // brick-major (kfb_common.cuh: vol_index)
    int X, Y, Z;      // global dims
    int bx, by, bz0;  // brick grid of the stored planes
    int zb, ze;       // planes whose voxels this context owns: [zb, ze)
    float vs[3];
    Pose aff;
    float *out;       // xyz triples
    unsigned long long cap;
    unsigned long long *counter;
};

__device__ __forceinline__ void unpack(uint32_t w, float &F, int &W)
{
    F = __fmul_rn((float)(short)(w & 0xffffu), KFB_DIVSHORTMAX);
    W = (int)(short)(w >> 16);
}

__global__ void __launch_bounds__(256) extract_kernel(const ExtractArgs a)
{
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const bool inside = x < a.X && y < a.Y;
    const float Vx = __fmul_rn(__fadd_rn((float)x, 0.5f), a.vs[0]);
    const float Vy = __fmul_rn(__fadd_rn((float)y, 0.5f), a.vs[1]);
    const int zlast = min(a.ze, a.Z - 1); // z + 1 must exist (tsdf_volume.cu:340)
    uint32_t wz = 0;
    if (inside && a.zb < zlast) wz = __ldg(a.vol + vol_index(a.bx, a.by, a.bz0, x, y, a.zb));
    const unsigned lane = threadIdx.x; // blockDim.x == 32
    for (int z = a.zb; z < zlast; ++z)
    {
        float pts[9];
        int cnt = 0;
        uint32_t wnext = 0;
        if (inside)
        {
            wnext = __ldg(a.vol + vol_index(a.bx, a.by, a.bz0, x, y, z + 1));
            float F; int W;
            unpack(wz, F, W);
            if (W != 0 && F != 1.f)
            {
                const float Vz = __fmul_rn(__fadd_rn((float)z, 0.5f), a.vs[2]);
#pragma unroll
                for (int axis = 0; axis < 3; ++axis)
                {
                    uint32_t wn;
                    if (axis == 0) { if (x + 1 >= a.X) continue; wn = __ldg(a.vol + vol_index(a.bx, a.by, a.bz0, x + 1, y, z)); }
                    else if (axis == 1) { if (y + 1 >= a.Y) continue; wn = __ldg(a.vol + vol_index(a.bx, a.by, a.bz0, x, y + 1, z)); }
                    else wn = wnext;
                    float Fn; int Wn;
                    unpack(wn, Fn, Wn);
                    if (Wn == 0 || Fn == 1.f) continue;
                    if (!((F > 0.f && Fn < 0.f) || (F < 0.f && Fn > 0.f))) continue;
                    const float V = axis == 0 ? Vx : (axis == 1 ? Vy : Vz);
                    const float Vn = __fadd_rn(V, a.vs[axis]);
                    const float d_inv = __fdiv_rn(1.f, __fadd_rn(fabsf(F), fabsf(Fn)));
                    const float pi = __fmul_rn(__fmaf_rn(fabsf(F), Vn, __fmul_rn(V, fabsf(Fn))), d_inv);
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                    {
                        const float r0 = a.aff.R.m[3 * c], r1 = a.aff.R.m[3 * c + 1], r2 = a.aff.R.m[3 * c + 2];
                        float o;
                        if (axis == 0) o = __fmaf_rn(Vz, r2, __fmaf_rn(pi, r0, __fmul_rn(Vy, r1)));
                        else if (axis == 1) o = __fmaf_rn(Vz, r2, __fmaf_rn(pi, r1, __fmul_rn(Vx, r0)));
                        else o = __fmaf_rn(pi, r2, __fadd_rn(__fmul_rn(Vy, r1), __fmul_rn(Vx, r0)));
                        pts[3 * cnt + c] = __fadd_rn(o, a.aff.t[c]);
                    }
                    ++cnt;
                }
            }
        }
        wz = wnext;
        // warp-aggregated append
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (unsigned)o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total > 0)
        {
            unsigned long long base = 0;
            if (lane == 31) base = atomicAdd(a.counter, (unsigned long long)total);
            base = __shfl_sync(0xffffffffu, base, 31);
            unsigned long long pos = base + (unsigned long long)(incl - cnt);
            for (int i = 0; i < cnt; ++i, ++pos)
                if (pos < a.cap)
                {
                    a.out[3 * pos + 0] = pts[3 * i + 0];
                    a.out[3 * pos + 1] = pts[3 * i + 1];
                    a.out[3 * pos + 2] = pts[3 * i + 2];
                }
        }
    }
}

int launch_extract(kfb_ctx *ctx, const float volpose12[12], float *host_points3, size_t cap, size_t *n_points)
{
    if (cap > ctx->cloud_cap)
    {
        if (ctx->cloud) cudaFree(ctx->cloud);
        ctx->cloud = nullptr;
        ctx->cloud_cap = 0;
        KFB_CUDA(ctx, cudaMalloc(&ctx->cloud, cap * 3 * sizeof(float)));
        ctx->cloud_cap = cap;
    }
    ExtractArgs a;
    a.vol = ctx->vol;
    a.X = ctx->p.volu_dims[0]; a.Y = ctx->p.volu_dims[1]; a.Z = ctx->p.volu_dims[2];
    a.bx = ctx->bdim[0]; a.by = ctx->bdim[1]; a.bz0 = ctx->bz0;
    const bool slab = ctx->p.slab_z_end > ctx->p.slab_z_begin;
    a.zb = slab ? ctx->p.slab_z_begin : 0;
    a.ze = slab ? ctx->p.slab_z_end : a.Z;
    for (int i = 0; i < 3; ++i) a.vs[i] = ctx->voxel_size[i];
    a.aff = make_pose(volpose12);
    a.out = ctx->cloud;
    a.cap = cap;
    a.counter = ctx->counters + 1;
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->counters + 1, 0, sizeof(unsigned long long), ctx->stream));
    dim3 block(32, 8), grid((a.X + 31) / 32, (a.Y + 7) / 8);
    extract_kernel<<<grid, block, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    KFB_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host + 1, ctx->counters + 1, sizeof(unsigned long long),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    size_t n = (size_t)ctx->counters_host[1];
    if (n > cap) n = cap;
    if (n) KFB_CUDA(ctx, cudaMemcpy(host_points3, ctx->cloud, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    *n_points = n;
    return KFB_OK;
}

} // namespace kfb
