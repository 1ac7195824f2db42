// kfb_integrate.cu -- TSDF integration (replaces kf::device::integrate,
// kfusion/src/tsdf_volume.cu:34-111) and volume reset (tsdf_volume.cu:11-32).
//
// Voxel-column sweep over the packed {int16 tsdf, int16 weight} volume: a thread owns four consecutive x
// voxels and marches z; one 128-bit load and one 128-bit store per z step, issued only when at least one of
// the four voxels passes the reference's update predicate (culled voxels cost no HBM traffic).  A warp owns a
// compact 16 x 8 voxel patch, a block four of them, and the grid's z dimension cuts the sweep into chunks.
// The predicate and the update arithmetic are the reference's, rounding for rounding:
//   * vc(z) is the reference's running sum, vc = fma(voxel_size.x, R[:,2], vc), from z = 1; it is carried as
//     packed f32x2 pairs (FFMA2); column_states_kernel walks every column once and stores the sums at the
//     chunk starts, a far z-slab's prefix is crossed by an exact integer jump (jump4);
//   * pixel = round-half-even(fma(vc.x * MUFU.RCP(vc.z), fx, cx)), done with the 2^23 magic-number add so it
//     stays off the conversion pipe;
//   * sdf = depth - |vc| / lambda is only evaluated exactly inside a narrow band around the surface: a
//     per-pixel table of conservative vc.z thresholds classifies "certainly free space => tsdf == 1.0f exactly"
//     and "certainly behind the surface => rejected" with two compares.  The thresholds are proved
//     conservative in build_tables_kernel, so classification never changes a result, only skips work.
// Everything that is not the reference's arithmetic was taken off the per-voxel path (the first version was
// instruction-issue bound, profiles/r01_v1_*):
//   * a per-thread conservative frustum interval [za, zb] in z (vc(z) is affine in z, so the four image-border
//     inequalities, vc.z > 0 and vc.z <= max accepted depth are half-lines in z) and an occlusion cut from a
//     max-pyramid of the thresholds over the pixels the columns can reach; planes outside are not visited,
//     inside the interval the exact per-voxel predicate still decides;
//   * a deep-free-space path: while every pixel a warp's columns can land on still sees free space at a
//     plane's largest vc.z, all voxels of that plane get tsdf = 1.0f -- no projection, no running sums.  vc.z
//     grows with z, so these planes are a prefix of a chunk: the warp streams through it and takes the general
//     path only behind it;
//   * in the general path every stage's dependent loads (thresholds, exact depth, weight table, brick flags)
//     are issued together: the warps wait on memory with every slot of the SM taken, so time is the sum of the
//     warps' lifetimes;
//   * the update runs without the quarter-rate XU pipe: int->float by magic-number add, the reciprocal of
//     weight+1 from a device-built table of MUFU.RCP results, float->int truncation by an RZ add of 2^23;
//   * when a thread's four voxels hold the same word and receive the same tsdf (free space), the update is
//     computed once; stores whose value equals the loaded one are dropped.
// Measured (profiles/): 56 M warp instructions per 640x480 / 512^3 frame (the first version: 271 M), DRAM
// traffic within 10 % of the algorithmic bytes; 82 % of the measured HBM copy bandwidth when every voxel is updated.
// Side product for the raycaster: a voxel that turns negative marks the 8^3 bricks within two voxels of it in
// a byte map; three separable passes turn the map into the brick distance field kfb_raycast.cu skips with.
#include "kfb_common.cuh"
#include <algorithm>
#include <cmath>
#include <vector>
#include <cstdlib>

namespace kfb
{

struct CullPlane // conservative half-line in z: g0(x, y) + z * g1 >= 0
{
    float a, b, g;  // g0 = a*vc.x + b*vc.y + g*vc.z at z = 0
    float slack;    // added to g0 (float-evaluation and running-sum error bounds)
    float ninv;     // -1 / g1
    int kind;       // 0: z >= g0*ninv, 1: z <= g0*ninv, 2: constant (g0 < 0 => empty), 3: ignore
};
#define KFB_NCULL 6

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
    const float2 *thrz;
    unsigned long long *states; // [chunk][6][X/4 * Y]: packed vc of every thread after plane zstart(chunk) - 1
    int2 *col_range;            // [X/4 * Y]: first and last plane of the column group's frustum interval (empty: {1, 0})
    int nchunks;
    const float2 *exact;
    const float4 *wtab;
    const float *zexit;
    const float2 *zmip;         // pyramid of {max lo_z, min hi_z}, levels 2..7 (tiles of 4..128 px), see build_zmip_kernel
    int mip_off[6], mip_w[6];
    float Sx, Sy, Sz, invSz, driftE; // per-plane step of vc (float), 1/Sz, bound on the running-sum drift
    int max_weight;
    int no_fastpath;            // KFB_INTEGRATE_NOFAST: disable the deep-free-space path (tuning / testing)
    int no_prefix;              // KFB_INTEGRATE_NOPREFIX: fast path only for warps whose whole interval is free space
    int diag;                   // KFB_INTEGRATE_DIAG (timing experiments ONLY, results are wrong): 1 = general-path warps return, 2 = fast-path warps return
    int use_jump, jump_min; // exact jump of the running sum for prefixes of at least jump_min planes
    uint8_t *bricks;
    int *bdirty;         // set when a brick flag flips 0 -> 1 (the distance map must be rebuilt)
    int bx, by, bz, bz0; // brick grid dims (x, y, stored z bricks) and first stored z brick
    CullPlane cull[KFB_NCULL];
    unsigned long long *counter;
    // ---- work plan (integrate_plan_kernel): the sweep as two lists of (patch, plane range) items ----
    const float4 *tab4;           // per pixel {hi_z, lo_z, depth, 1/lambda}
    const float2 *zsparse;        // sparse table of {max lo_z, min hi_z}: level k (1..6) at pixel (x, y) covers the 2^k x 2^k window from (x, y)
    int use_sparse;
    int npx, npy;                 // 16 x 8 voxel patches in x and y
    int mask_words;               // 32-bit words of a patch's "chunk has a general item" mask
    uint2 *items_stream;          // {patch, z0 | z1 << 16}: every voxel of these planes gets tsdf = 1.0f
    uint2 *items_general;         // {patch, z0 | z1 << 16}: planes that need the per-voxel predicate (within one chunk)
    unsigned int *plan_counts;    // [0] stream items, [1] general items
    unsigned int *patch_mask;     // [patch][mask_words]
    unsigned int *slot_of;        // [patch][chunk] -> general item index
    unsigned long long *gstates;  // [general item][6][32]: packed vc of the item's 32 threads after plane zstart(chunk) - 1
    int gstate_cap;               // general items that have a state slot (the others replay their running sums)
    int refine;                   // KFB_INTEGRATE_REFINE: per-thread refinement of a general item's plane range (measured slower; kept for experiments)
};

#define KFB_MAGIC_F 12582912.0f   // 1.5 * 2^23
#define KFB_MAGIC_I 0x4B400000
#define KFB_SKIP (-4.0f)

// ---- packed f32x2 helpers (sm_100a FFMA2 / FMUL2 / FADD2: two IEEE-rounded ops per issue) ------------
// NOTE: ptxas contracts mul.rn.f32x2 followed by add.rn.f32x2 into one FFMA2 (checked in SASS); this file
// never feeds a packed mul into a packed add, only mul -> fma and fma -> add, which cannot be contracted.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// ---- per-pixel tables ---------------------------------------------------------------
// exact[p] = {depth, MUFU.RCP(lambda)} with lambda = sqrt(((u-cx)/fx)^2 + ((v-cy)/fy)^2 + 1)
// computed with the reference's operations (tsdf_volume.cu:65-68, device_utils.cuh:22-27).
// thrz[p]  = {hi_z, lo_z}:  vc.z <= hi_z  ==> the reference's tsdf is exactly 1.0f
//                           vc.z >  lo_z  ==> the reference rejects the voxel (sdf < -trunc)
// Proof sketch (all quantities positive, eps = 2^-24).  Step 1, in terms of d2 = |vc|^2 as the kernel
// computes it:
//   nrm = sqrt_rn(d2) in sqrt(d2)(1 +- eps); il = MUFU.RCP(lambda) in (1/lambda)(1 +- 2^-22);
//   nsdf = RN(il*nrm - depth).  With T = trunc(1+1e-5):
//   d2 <= hi2 := ((depth - T - 1e-6) lambda)^2 (1-8e-6) => il*nrm <= depth - T (reals) => -nsdf >= T
//        => MUFU.RCP(trunc) * (-nsdf) >= (1+1e-5)(1-2^-22)(1-eps) > 1 => fmin(1, .) == 1.
//   d2 >  lo2 := ((depth + T + 1e-6) lambda)^2 (1+8e-6) => il*nrm - depth > T > trunc => nsdf > trunc.
//   The 1e-6 absorbs the float rounding of depth -+ T (depth < 8 m).
// Step 2, from d2 to vc.z.  A voxel that lands on pixel (u, v) has round(fma(qx, fx, cx)) == u with
// qx = RN(MUFU.RCP(vc.z) * vc.x), so |vc.x / vc.z - (u-cx)/fx| <= hx := 0.501/fx (0.5 px of rounding
// plus < 0.001 px for the fma rounding and the 2^-21 relative error of rcp*mul), likewise hy.  Hence
//   vc.z^2 Lmin^2 <= d2_real <= vc.z^2 Lmax^2,  Lmax^2 = 1 + (|lx|+hx)^2 + (|ly|+hy)^2,
//                                              Lmin^2 = 1 + max(|lx|-hx,0)^2 + max(|ly|-hy,0)^2,
// and d2 (three float ops) is within (1 +- 4 eps) of d2_real.  So
//   vc.z <= hi_z := sqrt(hi2)/Lmax (1-1e-6) => d2 <= hi2,     vc.z > lo_z := sqrt(lo2)/Lmin (1+1e-6) => d2 > lo2.
// Invalid depth (<= 0 or NaN) => {-1, -1}: everything rejected, as the reference does (`depth <= 0`
// skip; NaN depth makes sdf NaN, which fails `sdf >= -trunc`).
// *zexit = max lo_z over the image: a voxel with vc.z above it is rejected whatever pixel it lands on.
__global__ void build_tables_kernel(const float *__restrict__ depth, int w, int h, float fx, float fy, float cx,
                                    float cy, float trunc, float2 *__restrict__ thrz, float2 *__restrict__ exact,
                                    float4 *__restrict__ tab4, float *__restrict__ zexit)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y * blockDim.y + threadIdx.y;
    float lo_z = -1.f;
    if (u < w && v < h)
    {
        const int p = v * w + u;
        const float d = depth[p];
        const float lx = __fmul_rn(rcp_fdividef(fx), __fsub_rn((float)u, cx));
        const float ly = __fmul_rn(rcp_fdividef(fy), __fsub_rn((float)v, cy));
        const float lam = __fsqrt_rn(__fadd_rn(__fmaf_rn(lx, lx, __fmul_rn(ly, ly)), 1.0f));
        const float il = rcp_fdividef(lam);
        exact[p] = make_float2(d, il);
        float hi_z = -1.f;
        if (d > 0.f)
        {
            const float hx = 0.501f / fx, hy = 0.501f / fy;
            const float ax = fabsf(lx), ay = fabsf(ly);
            const float lmax2 = 1.f + (ax + hx) * (ax + hx) + (ay + hy) * (ay + hy);
            const float mx = fmaxf(ax - hx, 0.f), my = fmaxf(ay - hy, 0.f);
            const float lmin2 = 1.f + mx * mx + my * my;
            const float T = trunc * 1.00001f;
            const float a = d - T - 1e-6f;
            if (a > 0.f)
            {
                const float b = a * lam;
                const float hi2 = b * b * (1.f - 8e-6f);
                hi_z = sqrtf(hi2 / lmax2) * (1.f - 1e-6f);
            }
            const float c = (d + T + 1e-6f) * lam;
            const float lo2 = c * c * (1.f + 8e-6f);
            lo_z = sqrtf(lo2 / lmin2) * (1.f + 1e-6f);
        }
        thrz[p] = make_float2(hi_z, lo_z);
        tab4[p] = make_float4(hi_z, lo_z, d, il); // the sweep's general path reads everything about a pixel with one load
    }
    // image-wide max of lo_z (positive floats order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo_z = fmaxf(lo_z, __shfl_xor_sync(0xffffffffu, lo_z, o));
    if (((threadIdx.y * blockDim.x + threadIdx.x) & 31) == 0 && lo_z > 0.f) atomicMax((int *)zexit, __float_as_int(lo_z));
}

// Pyramid over the image of {max lo_z, min hi_z}: level l (2..7) holds, per 2^l x 2^l pixel tile, the largest
// vc.z any pixel of the tile would still accept and the largest vc.z that is free space (tsdf == 1) for EVERY
// pixel of the tile (-1 as soon as one pixel has no valid depth).  One block builds all levels of a 128 x 128
// pixel region.
__device__ __forceinline__ float2 mm2(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fminf(a.y, b.y)); }
__global__ void __launch_bounds__(256) build_zmip_kernel(const float2 *__restrict__ thrz, int w, int h, float2 *__restrict__ mip,
                                                         int o2, int o3, int o4, int o5, int o6, int o7)
{
    __shared__ float2 s2[32][33], s3[16][17], s4[8][9], s5[4][5], s6[2][3];
    const int rx = blockIdx.x * 128, ry = blockIdx.y * 128, t = threadIdx.x;
    const int w2 = (w + 3) >> 2, w3 = (w + 7) >> 3, w4 = (w + 15) >> 4, w5 = (w + 31) >> 5, w6 = (w + 63) >> 6, w7 = (w + 127) >> 7;
    const float2 none = make_float2(-1.f, 3.0e38f); // neutral element (tiles outside the image)
    for (int i = t; i < 1024; i += 256)
    {
        const int ty = i >> 5, tx = i & 31;
        float2 m = none;
        for (int dy = 0; dy < 4; ++dy)
            for (int dx = 0; dx < 4; ++dx)
            {
                const int x = rx + tx * 4 + dx, y = ry + ty * 4 + dy;
                if (x < w && y < h)
                {
                    const float2 th = thrz[(size_t)y * w + x]; // {hi_z, lo_z}
                    m = mm2(m, make_float2(th.y, th.x));
                }
            }
        s2[ty][tx] = m;
        const int gx = (rx >> 2) + tx, gy = (ry >> 2) + ty;
        if (gx < w2 && gy < ((h + 3) >> 2)) mip[o2 + gy * w2 + gx] = m;
    }
    __syncthreads();
    {
        const int ty = t >> 4, tx = t & 15;
        const float2 m = mm2(mm2(s2[2 * ty][2 * tx], s2[2 * ty][2 * tx + 1]), mm2(s2[2 * ty + 1][2 * tx], s2[2 * ty + 1][2 * tx + 1]));
        s3[ty][tx] = m;
        const int gx = (rx >> 3) + tx, gy = (ry >> 3) + ty;
        if (gx < w3 && gy < ((h + 7) >> 3)) mip[o3 + gy * w3 + gx] = m;
    }
    __syncthreads();
    if (t < 64)
    {
        const int ty = t >> 3, tx = t & 7;
        const float2 m = mm2(mm2(s3[2 * ty][2 * tx], s3[2 * ty][2 * tx + 1]), mm2(s3[2 * ty + 1][2 * tx], s3[2 * ty + 1][2 * tx + 1]));
        s4[ty][tx] = m;
        const int gx = (rx >> 4) + tx, gy = (ry >> 4) + ty;
        if (gx < w4 && gy < ((h + 15) >> 4)) mip[o4 + gy * w4 + gx] = m;
    }
    __syncthreads();
    if (t < 16)
    {
        const int ty = t >> 2, tx = t & 3;
        const float2 m = mm2(mm2(s4[2 * ty][2 * tx], s4[2 * ty][2 * tx + 1]), mm2(s4[2 * ty + 1][2 * tx], s4[2 * ty + 1][2 * tx + 1]));
        s5[ty][tx] = m;
        const int gx = (rx >> 5) + tx, gy = (ry >> 5) + ty;
        if (gx < w5 && gy < ((h + 31) >> 5)) mip[o5 + gy * w5 + gx] = m;
    }
    __syncthreads();
    if (t < 4)
    {
        const int ty = t >> 1, tx = t & 1;
        const float2 m = mm2(mm2(s5[2 * ty][2 * tx], s5[2 * ty][2 * tx + 1]), mm2(s5[2 * ty + 1][2 * tx], s5[2 * ty + 1][2 * tx + 1]));
        s6[ty][tx] = m;
        const int gx = (rx >> 6) + tx, gy = (ry >> 6) + ty;
        if (gx < w6 && gy < ((h + 63) >> 6)) mip[o6 + gy * w6 + gx] = m;
    }
    __syncthreads();
    if (t == 0) mip[o7 + blockIdx.y * w7 + blockIdx.x] = mm2(mm2(s6[0][0], s6[0][1]), mm2(s6[1][0], s6[1][1]));
}

// Sparse table over the image of {max lo_z, min hi_z}: level k holds, at EVERY pixel (x, y), the extremes over the
// 2^k x 2^k window that starts there (clipped to the image).  A pixel rectangle of any position and size is then
// covered exactly by a few overlapping windows -- no tile alignment, no slop: the plan's free-space prefix and
// occlusion cut see precisely the pixels a patch can land on (the tile pyramid above widens the rectangle to whole
// tiles, by up to half its size again; on surfaces seen at a grazing angle every extra pixel costs planes).
// Level k from level k - 1 (h = 2^(k-1)); level 1 straight from the per-pixel thresholds.
__global__ void __launch_bounds__(256) build_zsparse_kernel(const float2 *__restrict__ thrz, const float2 *__restrict__ prev, float2 *__restrict__ cur, int w, int h,
                                                            int half)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= w || y >= h) return;
    float2 m = make_float2(-1.f, 3.0e38f);
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i)
        {
            const int xx = x + i * half, yy = y + j * half;
            if (xx < w && yy < h)
            {
                if (prev) m = mm2(m, __ldg(prev + (size_t)yy * w + xx));
                else { const float2 th = __ldg(thrz + (size_t)yy * w + xx); m = mm2(m, make_float2(th.y, th.x)); } // {hi_z, lo_z} -> {lo_z, hi_z}
            }
        }
    cur[(size_t)y * w + x] = m;
}

// wtab[wt] = {(float)wt, MUFU.RCP(wt + 1), bits(min(wt + 1, max_weight) << 16), 0}: the weight-dependent
// operands of the running mean (tsdf_volume.cu:72-77), produced by the same device instructions the
// update would execute.
__global__ void build_wtab_kernel(float4 *__restrict__ wtab, int max_weight)
{
    const int wt = blockIdx.x * blockDim.x + threadIdx.x;
    if (wt > max_weight) return;
    const int wp1 = wt + 1;
    wtab[wt] = make_float4((float)wt, rcp_fdividef((float)wp1), __uint_as_float((unsigned)min(wp1, max_weight) << 16), 0.f);
}

// (the exact jump of the reference's running sums, jump_fma<N>, lives in kfb_common.cuh: the raycaster uses it too)
__device__ __forceinline__ void jump4(float v[4], float a, float b, int n) { jump_fma<4>(v, a, b, n); }

// ---- the update (tsdf_volume.cu:69-79) ---------------------------------------------------
// generic form, any stored word
__device__ __noinline__ unsigned int update_generic(unsigned int wv, float t, int max_weight)
{
    const int tsv = (int)(short)(wv & 0xffffu);
    const int wt = (int)(short)(wv >> 16);
    const float pre = __fmul_rn((float)tsv, KFB_DIVSHORTMAX);
    const int wp1 = wt + 1;
    const float rd = rcp_fdividef((float)wp1);
    const float nt = __fmul_rn(rd, __fmaf_rn(pre, (float)wt, t));
    int q = __float2int_rz(__fmul_rn(nt, (float)KFB_SHORTMAX));
    q = max(-KFB_SHORTMAX, min(KFB_SHORTMAX, q));
    const int nw = min(wp1, max_weight);
    return ((unsigned)q & 0xffffu) | ((unsigned)nw << 16);
}
// same result without the XU pipe for weights in [0, max_weight] (everything this library ever writes)
__device__ __forceinline__ unsigned int update_word_e(unsigned int wv, float t, const float4 e)
{
    const int tsv = (int)(short)(wv & 0xffffu);
    const float tsf = __fsub_rn(__int_as_float(KFB_MAGIC_I + tsv), KFB_MAGIC_F);     // (float)tsv, exact
    const float pre = __fmul_rn(tsf, KFB_DIVSHORTMAX);
    const float nt = __fmul_rn(e.y, __fmaf_rn(pre, e.x, t));
    const float s = __fmul_rn(nt, (float)KFB_SHORTMAX);                               // |s| < 2^16 here
    int qa = __float_as_int(__fadd_rz(fabsf(s), 8388608.0f)) - 0x4B000000;            // trunc(|s|)
    qa = min(qa, KFB_SHORTMAX);
    const int q = s < 0.f ? -qa : qa;
    return ((unsigned)q & 0xffffu) | __float_as_uint(e.z);
}
__device__ __forceinline__ unsigned int update_word(unsigned int wv, float t, const IntegrateArgs &a)
{
    const int wt = (int)wv >> 16;
    if ((unsigned)wt > (unsigned)a.max_weight) return update_generic(wv, t, a.max_weight);
    const float4 e = __ldg(a.wtab + wt);
    const int tsv = (int)(short)(wv & 0xffffu);
    const float tsf = __fsub_rn(__int_as_float(KFB_MAGIC_I + tsv), KFB_MAGIC_F);     // (float)tsv, exact
    const float pre = __fmul_rn(tsf, KFB_DIVSHORTMAX);
    const float nt = __fmul_rn(e.y, __fmaf_rn(pre, e.x, t));
    const float s = __fmul_rn(nt, (float)KFB_SHORTMAX);                               // |s| < 2^16 here
    int qa = __float_as_int(__fadd_rz(fabsf(s), 8388608.0f)) - 0x4B000000;            // trunc(|s|)
    qa = min(qa, KFB_SHORTMAX);
    const int q = s < 0.f ? -qa : qa;
    return ((unsigned)q & 0xffffu) | __float_as_uint(e.z);
}

// exact sdf evaluation for a voxel in the band around the surface (tsdf_volume.cu:63-71)
__device__ __forceinline__ float band_tsdf(const IntegrateArgs &a, unsigned long long xy, float cz, const float2 e, float rtrunc)
{
    float vxk, vyk;
    unpack2(xy, vxk, vyk);
    const float d2 = dot3c(vxk, vyk, cz, vxk, vyk, cz);
    const float nsdf = __fmaf_rn(e.y, __fsqrt_rn(d2), -e.x);
    return nsdf <= a.trunc ? fminf(1.f, __fmul_rn(rtrunc, -nsdf)) : KFB_SKIP;
}
// exact per-voxel predicate + tsdf for any vc.z (cameras inside the volume); cold path
__device__ __forceinline__ float classify_generic(const IntegrateArgs &a, float cx_, float cy_, float cz, float rtrunc)
{
    float t = KFB_SKIP;
    if (!(cz <= 0.f))
    {
        float qx, qy;
        if (cz >= KFB_FLT_MIN)
        {
            const float r = mufu_rcp(cz);
            qx = __fmul_rn(r, cx_);
            qy = __fmul_rn(r, cy_);
        }
        else
        {
            qx = __fdividef(cx_, cz);
            qy = __fdividef(cy_, cz);
        }
        const int ui = __float_as_int(__fadd_rn(__fmaf_rn(qx, a.fx, a.cx), KFB_MAGIC_F)) - KFB_MAGIC_I;
        const int vi = __float_as_int(__fadd_rn(__fmaf_rn(qy, a.fy, a.cy), KFB_MAGIC_F)) - KFB_MAGIC_I;
        if ((unsigned)ui < (unsigned)a.w && (unsigned)vi < (unsigned)a.h)
        {
            const int p = vi * a.w + ui;
            const float2 e = __ldg(a.exact + p);
            if (e.x > 0.f)
            {
                const float d2 = dot3c(cx_, cy_, cz, cx_, cy_, cz);
                const float nsdf = __fmaf_rn(e.y, __fsqrt_rn(d2), -e.x);
                if (nsdf <= a.trunc) t = fminf(1.f, __fmul_rn(rtrunc, -nsdf));
            }
        }
    }
    return t;
}

// bricks within two voxels of a sample that just turned negative (see kfb_raycast.cu for why two)
__device__ __noinline__ void mark_bricks(uint8_t *flags, int *dirty, int gbx, int gby, int gbz, int gbz0, int x0, int y, int z)
{
    // each axis reaches at most two bricks (x: 8 voxels from x0 - 2; y, z: +-2): the <= 8 flags are read together
    const int bx0 = max(x0 - 2, 0) >> 3, bx1 = min((x0 + 5) >> 3, gbx - 1);
    const int by0 = max(y - 2, 0) >> 3, by1 = min((y + 2) >> 3, gby - 1);
    const int bz0 = max((max(z - 2, 0) >> 3) - gbz0, 0), bz1 = min(((z + 2) >> 3) - gbz0, gbz - 1);
    if (bz1 < bz0 || by1 < by0 || bx1 < bx0) return;
    uint8_t *f[8];
    uint8_t v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        f[i] = flags + ((size_t)((i & 4) ? bz1 : bz0) * gby + ((i & 2) ? by1 : by0)) * gbx + ((i & 1) ? bx1 : bx0);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = *f[i];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (v[i] == 0) { *f[i] = 1; any = true; }
    if (any) *dirty = 1;
}

// Conservative interval [lo, hi] of planes in [zstart, zend) on which any of a thread's four columns can pass the
// reference's predicate (never excludes a voxel the exact predicate would accept: each plane is relaxed by `slack`
// and the bound is widened by one step).  (ax, ay, az) / (bx, by, bz) = vc of the first / last column at z = 0.
__device__ __forceinline__ void frustum_interval(const IntegrateArgs &a, float ax, float ay, float az, float bx, float by, float bz,
                                                 int zstart, int zend, float &lo, float &hi)
{
    lo = (float)zstart;
    hi = (float)(zend - 1);
    const float zx = __ldg(a.zexit);
#pragma unroll
    for (int c = 0; c < KFB_NCULL; ++c)
    {
        const CullPlane &cp = a.cull[c];
        if (cp.kind == 3) continue;
        const float ga = fmaf(cp.a, ax, fmaf(cp.b, ay, cp.g * az));
        const float gb = fmaf(cp.a, bx, fmaf(cp.b, by, cp.g * bz));
        float g0 = fmaxf(ga, gb) + cp.slack;
        if (c == KFB_NCULL - 1) g0 += zx; // vc.z <= zexit
        const float zc = g0 * cp.ninv;
        if (cp.kind == 0) lo = fmaxf(lo, zc - 1.f);
        else if (cp.kind == 1) hi = fminf(hi, zc + 1.f);
        else if (g0 < 0.f) hi = -1.f;
    }
    lo = fminf(lo, (float)zend);
    hi = fmaxf(hi, (float)zstart - 2.f);
}

// phase B of one plane: running weighted mean, re-encode, store (tsdf_volume.cu:69-79)
template <bool COUNT>
__device__ __forceinline__ void update_quad(const IntegrateArgs &a, uint4 *vp, const uint4 wd, const float t[4], int x0, int y, int z,
                                            unsigned int &n_upd)
{
    uint4 o = wd;
    const bool uni = (wd.x == wd.y) & (wd.x == wd.z) & (wd.x == wd.w) & (t[0] == t[1]) & (t[0] == t[2]) & (t[0] == t[3]);
    if (uni)
    {
        o.x = o.y = o.z = o.w = update_word(wd.x, t[0], a);
        if (COUNT) n_upd += 4;
    }
    else
    {
        const unsigned int w0 = wd.x >> 16, w1 = wd.y >> 16, w2 = wd.z >> 16, w3 = wd.w >> 16, mw = (unsigned)a.max_weight;
        if (max(max(w0, w1), max(w2, w3)) <= mw)
        {
            // the four table entries together (one latency), then branch-free selects
            const float4 e0 = __ldg(a.wtab + w0), e1 = __ldg(a.wtab + w1), e2 = __ldg(a.wtab + w2), e3 = __ldg(a.wtab + w3);
            const unsigned int n0 = update_word_e(wd.x, t[0], e0), n1 = update_word_e(wd.y, t[1], e1);
            const unsigned int n2 = update_word_e(wd.z, t[2], e2), n3 = update_word_e(wd.w, t[3], e3);
            o.x = t[0] != KFB_SKIP ? n0 : wd.x;
            o.y = t[1] != KFB_SKIP ? n1 : wd.y;
            o.z = t[2] != KFB_SKIP ? n2 : wd.z;
            o.w = t[3] != KFB_SKIP ? n3 : wd.w;
            if (COUNT) n_upd += (t[0] != KFB_SKIP) + (t[1] != KFB_SKIP) + (t[2] != KFB_SKIP) + (t[3] != KFB_SKIP);
        }
        else
        {
            if (t[0] != KFB_SKIP) { o.x = update_word(wd.x, t[0], a); if (COUNT) ++n_upd; }
            if (t[1] != KFB_SKIP) { o.y = update_word(wd.y, t[1], a); if (COUNT) ++n_upd; }
            if (t[2] != KFB_SKIP) { o.z = update_word(wd.z, t[2], a); if (COUNT) ++n_upd; }
            if (t[3] != KFB_SKIP) { o.w = update_word(wd.w, t[3], a); if (COUNT) ++n_upd; }
        }
    }
    if ((o.x != wd.x) | (o.y != wd.y) | (o.z != wd.z) | (o.w != wd.w))
    {
        __stcs(vp, o);
        // a voxel that turns negative here (it was not before) activates the bricks around it
        if (((o.x & ~wd.x) | (o.y & ~wd.y) | (o.z & ~wd.z) | (o.w & ~wd.w)) & 0x8000u)
            mark_bricks(a.bricks, a.bdirty, a.bx, a.by, a.bz, a.bz0, x0, y, z);
    }
}

// ---- the sweep ------------------------------------------------------------------------
#define KFB_BAND (-8.0f)
#ifndef KFB_INT_PX
#define KFB_INT_PX 4 // threads of a warp along x (each owns 4 voxels); 32 / KFB_INT_PX rows
#endif
#ifndef KFB_INT_WARPS
#define KFB_INT_WARPS 4 // warps per block (one warp per block was measured slower: 32k tiny blocks per chunk layer)
#endif
#ifndef KFB_INT_MINB
#define KFB_INT_MINB 8
#endif
template <int U, bool COUNT>
__global__ void __launch_bounds__(32 * KFB_INT_WARPS, KFB_INT_MINB * 4 / KFB_INT_WARPS) integrate_kernel(const IntegrateArgs a)
{
    // a warp owns a compact 16 x 8 voxel patch (KFB_INT_PX = 4 threads x 8 rows; 64 B per row): its columns see
    // nearly the same part of the image, so warp-level decisions (fast path, loop bounds) are mostly unanimous
    const int x0 = (blockIdx.x * (KFB_INT_WARPS * KFB_INT_PX) + threadIdx.y * KFB_INT_PX + (threadIdx.x & (KFB_INT_PX - 1))) * 4;
    const int y = blockIdx.y * (32 / KFB_INT_PX) + (threadIdx.x / KFB_INT_PX);
    if (x0 >= a.X || y >= a.Y) return;

    const int zstart = a.zb + blockIdx.z * a.zchunk;
    const int zend = min(zstart + a.zchunk, a.ze);
    if (zstart >= zend) return;

    // whole-column frustum interval, computed once per column by column_states_kernel (see frustum_interval)
    const int2 cr = __ldg(a.col_range + (size_t)y * (a.X >> 2) + (x0 >> 2));
    const int za = max(zstart, cr.x);
    int zb = min(zend - 1, cr.y);
    if (za > zb) return;
    // vc at z = 0 of the first and the last of the four columns: R * (x*vs.x, y*vs.y, 0*vs.z) + t
    // (tsdf_volume.cu:49-50); the running sums themselves come from the stored chunk states below
    unsigned long long xy[4], zz[2];
    float z0v[4];
    {
        const float py = __fmul_rn((float)y, a.vsy);
        const float pz = __fmul_rn(0.f, a.vsz);
#pragma unroll
        for (int k = 0; k < 4; k += 3)
        {
            const float px = __fmul_rn((float)(x0 + k), a.vsx);
            const float3 r = rot3(a.pose.R, px, py, pz);
            z0v[k] = __fadd_rn(r.z, a.pose.t[2]);
            xy[k] = pack2(__fadd_rn(r.x, a.pose.t[0]), __fadd_rn(r.y, a.pose.t[1]));
        }
    }
    // last plane of this thread's interval that certainly still is deep free space (za - 1: none)
    int free_end = za - 1;
    // Occlusion cut: over planes [za, zb] the four columns project into a pixel rectangle (a line segment per
    // column; computed from the affine model and widened by the drift bound).  A voxel is rejected once vc.z
    // exceeds lo_z of its pixel, hence certainly once it exceeds the maximum of lo_z over that rectangle, which
    // the max-pyramid gives with four lookups.  Planes beyond that are never visited.
    if (a.Sz > 1e-6f)
    {
        float ax, ay, bx_, by_;
        unpack2(xy[0], ax, ay);
        unpack2(xy[3], bx_, by_);
        float umin = 1e30f, umax = -1e30f, vmin = 1e30f, vmax = -1e30f, zmin = 1e30f;
#pragma unroll
        for (int e = 0; e < 2; ++e)
        {
            const float zf = (float)(e ? zb : za);
#pragma unroll
            for (int k = 0; k < 2; ++k)
            {
                const float X = fmaf(zf, a.Sx, k ? bx_ : ax), Y = fmaf(zf, a.Sy, k ? by_ : ay), Zc = fmaf(zf, a.Sz, k ? z0v[3] : z0v[0]);
                const float r = mufu_rcp(fmaxf(Zc, 1e-3f)); // approximate is fine: the rectangle is padded below
                const float u = fmaf(a.fx * X, r, a.cx), v = fmaf(a.fy * Y, r, a.cy);
                umin = fminf(umin, u); umax = fmaxf(umax, u);
                vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
                zmin = fminf(zmin, Zc);
            }
        }
        if (zmin > 0.05f)
        {
            // pixel error of the model: (fx + |u - cx|) * E / z per axis, plus rounding to the nearest pixel
            const float pad = 1.5f + (a.fx + a.fy + (float)(a.w + a.h)) * a.driftE * (1.001f * mufu_rcp(zmin));
            const bool all_inside = umin - pad >= 0.f && umax + pad <= (float)(a.w - 1) && vmin - pad >= 0.f && vmax + pad <= (float)(a.h - 1);
            const int u0 = max((int)floorf(fmaxf(umin - pad, -1e6f)), 0), u1 = min((int)ceilf(fminf(umax + pad, 1e6f)), a.w - 1);
            const int v0 = max((int)floorf(fmaxf(vmin - pad, -1e6f)), 0), v1 = min((int)ceilf(fminf(vmax + pad, 1e6f)), a.h - 1);
            if (u0 > u1 || v0 > v1) return; // never inside the image on these planes
            const int span = max(u1 - u0, v1 - v0) + 1;
            // tiles a quarter of the span wide: the rectangle is covered by at most 5 x 5 of them
            const int l = max(32 - __clz(span - 1) - 2, 2);
            if (l <= 7)
            {
                const float2 *m = a.zmip + a.mip_off[l - 2];
                const int mw = a.mip_w[l - 2];
                float2 q = make_float2(-1.f, 3.0e38f);
                for (int ty = v0 >> l; ty <= (v1 >> l); ++ty)
                    for (int tx = u0 >> l; tx <= (u1 >> l); ++tx) q = mm2(q, __ldg(m + ty * mw + tx));
                const float zc = (q.x + 2.f * a.driftE - fminf(z0v[0], z0v[3])) * a.invSz + 1.f;
                zb = min(zb, (int)ceilf(fminf(zc, 1e6f)));
                if (za > zb) return;
                // Deep free space: while the largest vc.z of a plane (+ drift) does not exceed the smallest hi_z
                // of the pixels the columns can land on (all inside the image), every voxel of that plane passes
                // the predicate with tsdf == 1.0f exactly; neither projection nor running sums are needed.
                // vc.z grows with z (Sz > 0), so these planes are a prefix [za, free_end] of the interval.
                if (all_inside && !a.no_fastpath)
                {
                    const float zmaxv = fmaxf(z0v[0], z0v[3]), e2 = 2.f * a.driftE;
                    int zf = min(zb, (int)floorf(fminf(fmaxf((q.y - e2 - zmaxv) * a.invSz, -1.f), 1e6f)));
                    // the estimate may be off by rounding: step back until the exact test holds
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        if (zf >= za && fmaf((float)zf, a.Sz, zmaxv) + e2 > q.y) --zf;
                    if (zf >= za && fmaf((float)zf, a.Sz, zmaxv) + e2 <= q.y) free_end = zf;
                }
            }
        }
    }
    // The fast prefix only pays for planes on which the WHOLE warp takes it (a mixed plane would run both paths
    // in turn): planes up to the warp's smallest free_end.  Whatever set of lanes votes together, a lane's own
    // free_end is >= the minimum it receives, so the split can never change a result.
    int zfw;
    {
        const unsigned act = __activemask();
        zfw = __reduce_min_sync(act, free_end);
        if (a.no_prefix) zfw = __ballot_sync(act, free_end >= zb) == act ? 0x7fffffff : -0x7fffffff; // all or nothing
    }
    unsigned int n_upd = 0;
    const size_t plane4 = ((size_t)a.X * a.Y) >> 2; // uint4 per plane
    if (!(a.diag & 2))
    {
        const int pe = min(zfw, zb);
        uint4 *vp = reinterpret_cast<uint4 *>(a.vol) + ((size_t)(za - a.z_store0) * a.Y + y) * (a.X >> 2) + (x0 >> 2);
        const float ones[4] = {1.f, 1.f, 1.f, 1.f};
        int z = za;
        for (; z + 3 <= pe; z += 4, vp += 4 * plane4)
        {
            const uint4 w0 = __ldcs(vp), w1 = __ldcs(vp + plane4), w2 = __ldcs(vp + 2 * plane4), w3 = __ldcs(vp + 3 * plane4);
            update_quad<COUNT>(a, vp, w0, ones, x0, y, z, n_upd);
            update_quad<COUNT>(a, vp + plane4, w1, ones, x0, y, z + 1, n_upd);
            update_quad<COUNT>(a, vp + 2 * plane4, w2, ones, x0, y, z + 2, n_upd);
            update_quad<COUNT>(a, vp + 3 * plane4, w3, ones, x0, y, z + 3, n_upd);
        }
        for (; z <= pe; ++z, vp += plane4) update_quad<COUNT>(a, vp, __ldcs(vp), ones, x0, y, z, n_upd);
    }
    if (zfw >= zb || (a.diag & 1))
    {
        if (COUNT && n_upd) atomicAdd(a.counter, (unsigned long long)n_upd);
        return;
    }
    const int za_g = max(za, zfw + 1); // first plane of the general path

    const float sz = a.pose.R.m[8];
    const unsigned long long vs2 = pack2(a.vsx, a.vsx), sxy = pack2(a.pose.R.m[2], a.pose.R.m[5]), szz = pack2(sz, sz);
    // the reference's running sum (tsdf_volume.cu:56) up to the first visited plane: column_states_kernel has
    // stored vc after plane zstart - 1 for every chunk; only the planes zstart .. za - 1 are replayed here
    {
        const size_t nthr = (size_t)(a.X >> 2) * a.Y;
        const unsigned long long *st = a.states + (size_t)blockIdx.z * 6 * nthr + (size_t)y * (a.X >> 2) + (x0 >> 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) xy[k] = __ldg(st + (size_t)k * nthr);
        zz[0] = __ldg(st + 4 * nthr);
        zz[1] = __ldg(st + 5 * nthr);
#pragma unroll 4
        for (int z = zstart; z < za_g; ++z)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k) xy[k] = ffma2(vs2, sxy, xy[k]);
            zz[0] = ffma2(vs2, szz, zz[0]);
            zz[1] = ffma2(vs2, szz, zz[1]);
        }
    }
    // fast path needs vc.z >= FLT_MIN on every visited plane; vc.z is affine in z up to the running-sum
    // drift (millimetres at most), so the two ends decide with a 1 cm margin
    bool fast;
    {
        const float span = (float)(zb - za_g + 1) * __fmul_rn(a.vsx, sz);
        float c0, c1, c2, c3;
        unpack2(zz[0], c0, c1);
        unpack2(zz[1], c2, c3);
        const float m = fminf(fminf(c0, c1), fminf(c2, c3)); // plane za - 1
        fast = fminf(m, m + span) > 0.01f;
    }

    const float rtrunc = rcp_fdividef(a.trunc);
    uint4 *vp = reinterpret_cast<uint4 *>(a.vol) + ((size_t)(za_g - a.z_store0) * a.Y + y) * (a.X >> 2) + (x0 >> 2);

    if (!fast)
    {
        // cold path (camera within a centimetre of this column's planes): one plane at a time, any vc.z
        for (int z = za_g; z <= zb; ++z, vp += plane4)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k) xy[k] = ffma2(vs2, sxy, xy[k]);
            zz[0] = ffma2(vs2, szz, zz[0]);
            zz[1] = ffma2(vs2, szz, zz[1]);
            float cz[4], t[4];
            unpack2(zz[0], cz[0], cz[1]);
            unpack2(zz[1], cz[2], cz[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                float vxk, vyk;
                unpack2(xy[k], vxk, vyk);
                t[k] = classify_generic(a, vxk, vyk, cz[k], rtrunc);
            }
            if ((t[0] != KFB_SKIP) | (t[1] != KFB_SKIP) | (t[2] != KFB_SKIP) | (t[3] != KFB_SKIP))
                update_quad<COUNT>(a, vp, __ldcs(vp), t, x0, y, z, n_upd);
        }
        if (COUNT && n_upd) atomicAdd(a.counter, (unsigned long long)n_upd);
        return;
    }

    const unsigned long long fxy = pack2(a.fx, a.fy), cxy = pack2(a.cx, a.cy), magic2 = pack2(KFB_MAGIC_F, KFB_MAGIC_F);
    for (int z = za_g; z <= zb; z += U)
    {
        float ts[U][4], cz[U][4];
        unsigned int pix[U][4];
        float2 th[U][4];
        unsigned long long sxyv[U][4];
        const unsigned int last_pix = (unsigned int)(a.w * a.h - 1);
        // ---- phase A1: advance, project, issue all threshold loads -------------------------------
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k) xy[k] = ffma2(vs2, sxy, xy[k]);
            zz[0] = ffma2(vs2, szz, zz[0]);
            zz[1] = ffma2(vs2, szz, zz[1]);
            unpack2(zz[0], cz[u][0], cz[u][1]);
            unpack2(zz[1], cz[u][2], cz[u][3]);
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                sxyv[u][k] = xy[k];
                const float r = mufu_rcp(cz[u][k]);
                const unsigned long long q = fmul2(pack2(r, r), xy[k]);
                const unsigned long long m = fadd2(ffma2(q, fxy, cxy), magic2);
                float mu, mv;
                unpack2(m, mu, mv);
                const int ui = __float_as_int(mu) - KFB_MAGIC_I;
                const int vi = __float_as_int(mv) - KFB_MAGIC_I;
                const bool ok = ((unsigned)ui < (unsigned)a.w) & ((unsigned)vi < (unsigned)a.h);
                // an out-of-image voxel is classified as "behind everything": vc.z = +inf fails `<= hi_z` and
                // passes `> lo_z` for whatever (clamped) table entry it reads
                cz[u][k] = ok ? cz[u][k] : __int_as_float(0x7f800000);
                pix[u][k] = min((unsigned int)(vi * a.w + ui), last_pix);
                th[u][k] = __ldg(a.thrz + pix[u][k]);
            }
        }
        // ---- phase A2: classify against the thresholds; exact sdf only in the band -------------------
        bool band = false;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < 4; ++k)
            {
                const float t = cz[u][k] <= th[u][k].x ? 1.0f : (cz[u][k] > th[u][k].y ? KFB_SKIP : KFB_BAND);
                band |= (t == KFB_BAND);
                ts[u][k] = t;
            }
        if (band)
        {
            // all exact-depth entries first (the addresses are valid for every voxel), so that the warp waits for
            // one load latency and not for one per voxel a lane happens to have in the band
            float2 ex[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k) ex[u][k] = __ldg(a.exact + pix[u][k]);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const float tb = band_tsdf(a, sxyv[u][k], cz[u][k], ex[u][k], rtrunc);
                    ts[u][k] = ts[u][k] == KFB_BAND ? tb : ts[u][k];
                }
        }
        // ---- loads, then phase B ------------------------------------------------------------------------
        uint4 word[U];
        bool need[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
            need[u] = ((ts[u][0] != KFB_SKIP) | (ts[u][1] != KFB_SKIP) | (ts[u][2] != KFB_SKIP) | (ts[u][3] != KFB_SKIP)) & (z + u <= zb);
            if (need[u]) word[u] = __ldcs(vp + (size_t)u * plane4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (need[u]) update_quad<COUNT>(a, vp + (size_t)u * plane4, word[u], ts[u], x0, y, z + u, n_upd);
        vp += (size_t)U * plane4;
    }
    if (COUNT && n_upd) atomicAdd(a.counter, (unsigned long long)n_upd); // threads leave at different times: no shuffles
}

// vc of every thread's four columns at the start of every z-chunk, by the reference's recurrence from z = 1
// (exact jump for a long prefix in front of a far z-slab).  One sequential pass per column instead of one
// replay per chunk.
__global__ void __launch_bounds__(128) column_states_kernel(const IntegrateArgs a)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x0 >= a.X || y >= a.Y) return;
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
    // columns that never enter the frustum need no states (their sweep threads return before reading them), and
    // no chunk past the last visited plane does
    int z_first, z_last;
    {
        float lo, hi;
        frustum_interval(a, vx[0], vy[0], vz[0], vx[3], vy[3], vz[3], a.zb, a.ze, lo, hi);
        z_first = max(a.zb, (int)floorf(lo));
        z_last = min(a.ze - 1, (int)ceilf(hi));
        a.col_range[(size_t)y * (a.X >> 2) + (x0 >> 2)] = make_int2(z_first, z_last);
        if (z_first > z_last) return;
    }
    const float sx = a.pose.R.m[2], sy = a.pose.R.m[5], sz = a.pose.R.m[8];
    int done = 0; // planes applied so far
    if (a.use_jump && a.zb - 1 >= a.jump_min)
    {
        jump4(vx, a.vsx, sx, a.zb - 1);
        jump4(vy, a.vsx, sy, a.zb - 1);
        jump4(vz, a.vsx, sz, a.zb - 1);
        done = a.zb - 1;
    }
    unsigned long long xy[4], zz[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) xy[k] = pack2(vx[k], vy[k]);
    zz[0] = pack2(vz[0], vz[1]);
    zz[1] = pack2(vz[2], vz[3]);
    const unsigned long long vs2 = pack2(a.vsx, a.vsx), sxy = pack2(sx, sy), szz = pack2(sz, sz);
    const size_t nthr = (size_t)(a.X >> 2) * a.Y;
    unsigned long long *st = a.states + (size_t)y * (a.X >> 2) + (x0 >> 2);
    for (int c = 0; c < a.nchunks; ++c)
    {
        const int target = a.zb + c * a.zchunk - 1; // state after this plane
        if (target + 1 > z_last) break;
#pragma unroll 4
        for (; done < target; ++done)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k) xy[k] = ffma2(vs2, sxy, xy[k]);
            zz[0] = ffma2(vs2, szz, zz[0]);
            zz[1] = ffma2(vs2, szz, zz[1]);
        }
        if (target + a.zchunk < z_first) continue; // the chunk ends before the first visited plane: nobody reads its state
        unsigned long long *o = st + (size_t)c * 6 * nthr;
#pragma unroll
        for (int k = 0; k < 4; ++k) o[(size_t)k * nthr] = xy[k];
        o[4 * nthr] = zz[0];
        o[5 * nthr] = zz[1];
    }
}

// =====================================================================================================
// The sweep as a work plan (round 2).  The kernel above decides per warp and per z-chunk, inside the sweep, what
// its patch needs (frustum interval, occlusion cut, deep free space) before it touches a voxel; with every warp
// slot taken by warps that wait on memory, that setup, the streaming part and the per-voxel part simply add up
// (DESIGN.md 3.2).  Here the decisions are taken once, by a small kernel with one thread per (16 x 8 voxel patch,
// 16-plane chunk), which writes two compact work lists:
//   stream items  {patch, z0..z1}: every voxel of these planes passes the reference's predicate with tsdf == 1.0f
//                 exactly (same proof as the fast path above, at patch granularity) -- integrate_stream_kernel does
//                 nothing but load -> running mean -> store, eight planes in flight per thread;
//   general items {patch, z0..z1}: planes on which the exact per-voxel predicate decides (the band around the
//                 surface, the frustum border, holes) -- integrate_general_kernel runs the reference's arithmetic
//                 as a two-stage software pipeline: the table entries and the voxel words of plane z + 1 are in
//                 flight while plane z is classified and updated, so a warp pays one memory latency per plane
//                 instead of three dependent ones.
// The running sums of the reference (vc += zstep) are only needed by the general items: integrate_states_kernel
// walks the patches that have any, one warp per patch, and stores the 32 threads' sums at the chunk starts the
// plan asked for (48 B per thread per general item instead of per column per chunk).
// Culling is conservative and the per-voxel predicate is exact, so no split of the work can change a result:
// tests/test_ref_ab.py and tests/test_ref_full.py compare whole volumes with the reference kernels' bit for bit.
#define KFB_PATCH_X 16
#define KFB_PATCH_Y 8
#define KFB_PLAN_ZCHUNK 8

// conservative interval of planes on which any column of a patch can pass the predicate; (cx, cy, cz)[k] = vc at
// z = 0 of the patch's corner columns (g is affine in x and y, so its maximum over the patch sits at a corner)
template <int N>
__device__ __forceinline__ void frustum_interval_n(const IntegrateArgs &a, const float cx[N], const float cy[N], const float cz[N], int zstart,
                                                   int zend, float &lo, float &hi)
{
    lo = (float)zstart;
    hi = (float)(zend - 1);
    const float zx = __ldg(a.zexit);
#pragma unroll
    for (int c = 0; c < KFB_NCULL; ++c)
    {
        const CullPlane &cp = a.cull[c];
        if (cp.kind == 3) continue;
        float g0 = -3.0e38f;
#pragma unroll
        for (int k = 0; k < N; ++k) g0 = fmaxf(g0, fmaf(cp.a, cx[k], fmaf(cp.b, cy[k], cp.g * cz[k])));
        g0 += cp.slack;
        if (c == KFB_NCULL - 1) g0 += zx; // vc.z <= zexit
        const float zc = g0 * cp.ninv;
        if (cp.kind == 0) lo = fmaxf(lo, zc - 1.f);
        else if (cp.kind == 1) hi = fminf(hi, zc + 1.f);
        else if (g0 < 0.f) hi = -1.f;
    }
    lo = fminf(lo, (float)zend);
    hi = fmaxf(hi, (float)zstart - 2.f);
}

__global__ void __launch_bounds__(128) integrate_plan_kernel(const IntegrateArgs a)
{
    const unsigned FULL = 0xffffffffu;
    const int t = blockIdx.x * blockDim.x + threadIdx.x; // (chunk, patch row, patch column), column fastest:
    const int npatch = a.npx * a.npy;                    // a warp plans a row of patches, its items stay neighbours
    bool s_item = false, g_item = false;
    int s_z0 = 0, s_z1 = 0, g_z0 = 0, g_z1 = 0, patch = 0, c = 0;
    if (t < npatch * a.nchunks)
    {
        c = t / npatch;
        patch = t - c * npatch;
        const int py = patch / a.npx, px = patch - py * a.npx;
        const int zstart = a.zb + c * a.zchunk, zend = min(zstart + a.zchunk, a.ze);
        const int xa = px * KFB_PATCH_X, xb = min(xa + KFB_PATCH_X - 1, a.X - 1), ya = py * KFB_PATCH_Y, yb = min(ya + KFB_PATCH_Y - 1, a.Y - 1);
        float cx[4], cy[4], cz[4];
        const float pz = __fmul_rn(0.f, a.vsz);
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const float3 r = rot3(a.pose.R, __fmul_rn((float)((k & 1) ? xb : xa), a.vsx), __fmul_rn((float)((k & 2) ? yb : ya), a.vsy), pz);
            cx[k] = __fadd_rn(r.x, a.pose.t[0]); cy[k] = __fadd_rn(r.y, a.pose.t[1]); cz[k] = __fadd_rn(r.z, a.pose.t[2]);
        }
        float lo, hi;
        frustum_interval_n<4>(a, cx, cy, cz, zstart, zend, lo, hi);
        int za = max(zstart, (int)floorf(lo)), zb = min(zend - 1, (int)ceilf(hi));
        int free_end = za - 1;
        // Occlusion cut and deep free space at patch granularity (see the per-thread version above for the proofs):
        // over planes [za, zb] the patch projects into a pixel rectangle; beyond max lo_z of it every voxel is
        // rejected, up to min hi_z of it (all pixels inside the image) every voxel is free space.
        if (za <= zb && a.Sz > 1e-6f)
        {
            float umin = 1e30f, umax = -1e30f, vmin = 1e30f, vmax = -1e30f, zmin = 1e30f;
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
                const float zf = (float)(e ? zb : za);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const float X = fmaf(zf, a.Sx, cx[k]), Y = fmaf(zf, a.Sy, cy[k]), Zc = fmaf(zf, a.Sz, cz[k]);
                    const float r = mufu_rcp(fmaxf(Zc, 1e-3f));
                    const float u = fmaf(a.fx * X, r, a.cx), v = fmaf(a.fy * Y, r, a.cy);
                    umin = fminf(umin, u); umax = fmaxf(umax, u);
                    vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
                    zmin = fminf(zmin, Zc);
                }
            }
            if (zmin > 0.05f)
            {
                const float pad = 1.5f + (a.fx + a.fy + (float)(a.w + a.h)) * a.driftE * (1.001f * mufu_rcp(zmin));
                const bool all_inside = umin - pad >= 0.f && umax + pad <= (float)(a.w - 1) && vmin - pad >= 0.f && vmax + pad <= (float)(a.h - 1);
                const int u0 = max((int)floorf(fmaxf(umin - pad, -1e6f)), 0), u1 = min((int)ceilf(fminf(umax + pad, 1e6f)), a.w - 1);
                const int v0 = max((int)floorf(fmaxf(vmin - pad, -1e6f)), 0), v1 = min((int)ceilf(fminf(vmax + pad, 1e6f)), a.h - 1);
                if (u0 > u1 || v0 > v1) zb = za - 1; // never inside the image on these planes
                else
                {
                    const int span = max(u1 - u0, v1 - v0) + 1;
                    const int l = max(32 - __clz(span - 1) - 2, 2); // tiles a quarter of the span wide: at most 5 x 5 cover the rectangle
                    if (l <= 7)
                    {
                        float2 q = make_float2(-1.f, 3.0e38f);
                        if (a.use_sparse)
                        {
                            // exact cover of the rectangle by overlapping 2^k x 2^k windows of the sparse table: k from
                            // the shorter side, raised until five windows reach along the longer one
                            const int W = u1 - u0 + 1, H = v1 - v0 + 1;
                            int k = min(max(31 - __clz(min(W, H)), 1), 6);
                            while (k < 6 && ((max(W, H) + (1 << k) - 1) >> k) > 5) ++k;
                            const int sw = 1 << k;
                            if (((max(W, H) + sw - 1) >> k) > 5) q = make_float2(3.0e38f, -1.f); // wider than 5 x 64 pixels: decide per voxel
                            else
                            {
                                const float2 *m = a.zsparse + (size_t)(k - 1) * a.w * a.h;
                                const int xl = max(u0, u1 - sw + 1), yl = max(v0, v1 - sw + 1), ny = (H + sw - 1) >> k;
                                for (int j = 0; j < ny; ++j)
                                {
                                    const int yy = min(v0 + j * sw, yl);
                                    float2 tl[5];
#pragma unroll
                                    for (int i = 0; i < 5; ++i) tl[i] = __ldg(m + (size_t)yy * a.w + min(u0 + i * sw, xl));
#pragma unroll
                                    for (int i = 0; i < 5; ++i) q = mm2(q, tl[i]);
                                }
                            }
                        }
                        else
                        {
                            const float2 *m = a.zmip + a.mip_off[l - 2];
                            const int mw = a.mip_w[l - 2];
                            const int tx0 = u0 >> l, tx1 = u1 >> l, ty0 = v0 >> l, ty1 = v1 >> l;
                            // 25 loads in flight; tiles beyond the rectangle repeat its last one (harmless for min / max)
                            float2 tl[25];
#pragma unroll
                            for (int j = 0; j < 5; ++j)
#pragma unroll
                                for (int i = 0; i < 5; ++i) tl[j * 5 + i] = __ldg(m + min(ty0 + j, ty1) * mw + min(tx0 + i, tx1));
#pragma unroll
                            for (int i = 0; i < 25; ++i) q = mm2(q, tl[i]);
                        }
                        const float zmin0 = fminf(fminf(cz[0], cz[1]), fminf(cz[2], cz[3])), zmax0 = fmaxf(fmaxf(cz[0], cz[1]), fmaxf(cz[2], cz[3]));
                        const float e2 = 2.f * a.driftE;
                        const float zc = (q.x + e2 - zmin0) * a.invSz + 1.f;
                        zb = min(zb, (int)ceilf(fminf(zc, 1e6f)));
                        if (all_inside && !a.no_fastpath && za <= zb)
                        {
                            int zf = min(zb, (int)floorf(fminf(fmaxf((q.y - e2 - zmax0) * a.invSz, -1.f), 1e6f)));
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                if (zf >= za && fmaf((float)zf, a.Sz, zmax0) + e2 > q.y) --zf;
                            if (zf >= za && fmaf((float)zf, a.Sz, zmax0) + e2 <= q.y) free_end = zf;
                        }
                    }
                }
            }
        }
        if (za <= zb)
        {
            const int fe = min(free_end, zb);
            if (fe >= za) { s_item = true; s_z0 = za; s_z1 = fe; }
            if (fe < zb) { g_item = true; g_z0 = max(za, fe + 1); g_z1 = zb; }
        }
    }
    // warp-aggregated append (keeps the warp's items in patch order)
    const unsigned ms = __ballot_sync(FULL, s_item), mg = __ballot_sync(FULL, g_item);
    const int lane = threadIdx.x & 31;
    unsigned bs = 0, bg = 0;
    if (lane == 0)
    {
        if (ms) bs = atomicAdd(a.plan_counts + 0, __popc(ms));
        if (mg) bg = atomicAdd(a.plan_counts + 1, __popc(mg));
    }
    bs = __shfl_sync(FULL, bs, 0);
    bg = __shfl_sync(FULL, bg, 0);
    const unsigned below = (1u << lane) - 1u;
    if (s_item) a.items_stream[bs + __popc(ms & below)] = make_uint2((unsigned)patch, (unsigned)s_z0 | ((unsigned)s_z1 << 16));
    if (g_item)
    {
        const unsigned idx = bg + __popc(mg & below);
        a.items_general[idx] = make_uint2((unsigned)patch, (unsigned)g_z0 | ((unsigned)g_z1 << 16));
        a.slot_of[(size_t)patch * a.nchunks + c] = idx;
        atomicOr(a.patch_mask + (size_t)patch * a.mask_words + (c >> 5), 1u << (c & 31));
    }
}

// running sums at the chunk starts of the general items: one warp per patch, lane = the sweep's thread of that patch
__global__ void __launch_bounds__(128) integrate_states_kernel(const IntegrateArgs a)
{
    const int patch = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (patch >= a.npx * a.npy) return;
    int c_last = -1;
    for (int w = a.mask_words - 1; w >= 0 && c_last < 0; --w)
    {
        const unsigned m = __ldg(a.patch_mask + (size_t)patch * a.mask_words + w);
        if (m) c_last = w * 32 + 31 - __clz(m);
    }
    if (c_last < 0) return;
    const int py = patch / a.npx, px = patch - py * a.npx;
    // lanes beyond the volume's edge compute a valid neighbour's sums (never read)
    const int x0 = min(px * KFB_PATCH_X + (lane & 3) * 4, a.X - 4), y = min(py * KFB_PATCH_Y + (lane >> 2), a.Y - 1);
    float vx[4], vy[4], vz[4];
    {
        const float pyf = __fmul_rn((float)y, a.vsy), pz = __fmul_rn(0.f, a.vsz);
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            const float3 r = rot3(a.pose.R, __fmul_rn((float)(x0 + k), a.vsx), pyf, pz);
            vx[k] = __fadd_rn(r.x, a.pose.t[0]); vy[k] = __fadd_rn(r.y, a.pose.t[1]); vz[k] = __fadd_rn(r.z, a.pose.t[2]);
        }
    }
    const float sx = a.pose.R.m[2], sy = a.pose.R.m[5], sz = a.pose.R.m[8];
    int done = 0; // planes applied so far
    if (a.use_jump && a.zb - 1 >= a.jump_min)
    {
        jump4(vx, a.vsx, sx, a.zb - 1);
        jump4(vy, a.vsx, sy, a.zb - 1);
        jump4(vz, a.vsx, sz, a.zb - 1);
        done = a.zb - 1;
    }
    unsigned long long xy[4], zz[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) xy[k] = pack2(vx[k], vy[k]);
    zz[0] = pack2(vz[0], vz[1]);
    zz[1] = pack2(vz[2], vz[3]);
    const unsigned long long vs2 = pack2(a.vsx, a.vsx), sxy = pack2(sx, sy), szz = pack2(sz, sz);
    for (int c = 0; c <= c_last; ++c)
    {
        const int target = a.zb + c * a.zchunk - 1; // state after this plane
#pragma unroll 4
        for (; done < target; ++done)
        {
#pragma unroll
            for (int k = 0; k < 4; ++k) xy[k] = ffma2(vs2, sxy, xy[k]);
            zz[0] = ffma2(vs2, szz, zz[0]);
            zz[1] = ffma2(vs2, szz, zz[1]);
        }
        if (!((__ldg(a.patch_mask + (size_t)patch * a.mask_words + (c >> 5)) >> (c & 31)) & 1u)) continue;
        const unsigned slot = __ldg(a.slot_of + (size_t)patch * a.nchunks + c);
        if (slot >= (unsigned)a.gstate_cap) continue; // no slot: the item replays by itself
        unsigned long long *o = a.gstates + (size_t)slot * 192 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k * 32] = xy[k];
        o[128] = zz[0];
        o[160] = zz[1];
    }
}

// running mean of a quad that receives tsdf = 1.0f in all four voxels; WT = per-weight table (shared or global)
template <bool COUNT>
__device__ __forceinline__ void update_free_quad(const IntegrateArgs &a, const float4 *__restrict__ wt, uint4 *vp, const uint4 wd, unsigned int &n_upd,
                                                 unsigned int &n_st)
{
    uint4 o;
    const unsigned int mw = (unsigned)a.max_weight;
    if ((wd.x == wd.y) & (wd.x == wd.z) & (wd.x == wd.w))
    {
        const unsigned int w0 = wd.x >> 16;
        o.x = w0 <= mw ? update_word_e(wd.x, 1.0f, wt[w0]) : update_generic(wd.x, 1.0f, a.max_weight);
        o.y = o.z = o.w = o.x;
    }
    else
    {
        const unsigned int w0 = wd.x >> 16, w1 = wd.y >> 16, w2 = wd.z >> 16, w3 = wd.w >> 16;
        if (max(max(w0, w1), max(w2, w3)) <= mw)
        {
            const float4 e0 = wt[w0], e1 = wt[w1], e2 = wt[w2], e3 = wt[w3];
            o.x = update_word_e(wd.x, 1.0f, e0); o.y = update_word_e(wd.y, 1.0f, e1);
            o.z = update_word_e(wd.z, 1.0f, e2); o.w = update_word_e(wd.w, 1.0f, e3);
        }
        else
        {
            o.x = update_generic(wd.x, 1.0f, a.max_weight); o.y = update_generic(wd.y, 1.0f, a.max_weight);
            o.z = update_generic(wd.z, 1.0f, a.max_weight); o.w = update_generic(wd.w, 1.0f, a.max_weight);
        }
    }
    if (COUNT) n_upd += 4;
    // a mean with +1 cannot turn a non-negative value negative: no brick can become active here
    if ((o.x != wd.x) | (o.y != wd.y) | (o.z != wd.z) | (o.w != wd.w))
    {
        __stcs(vp, o);
        if (COUNT) ++n_st;
    }
}

#define KFB_WTAB_SMEM 256 // per-weight table entries kept in shared memory (max_weight < 256; the default is 64)
#define KFB_STREAM_DEPTH 8
template <bool COUNT, bool SMEM>
__global__ void __launch_bounds__(128, 8) integrate_stream_kernel(const IntegrateArgs a)
{
    __shared__ float4 s_wt[SMEM ? KFB_WTAB_SMEM : 1];
    if (SMEM)
    {
        for (int i = threadIdx.x; i <= a.max_weight; i += blockDim.x) s_wt[i] = __ldg(a.wtab + i);
        __syncthreads();
    }
    const float4 *wt = SMEM ? s_wt : a.wtab;
    const int lane = threadIdx.x & 31;
    const unsigned int n_items = __ldg(a.plan_counts + 0);
    const size_t plane4 = ((size_t)a.X * a.Y) >> 2;
    unsigned int n_upd = 0, n_ld = 0, n_st = 0;
    for (unsigned int item = blockIdx.x * 4 + (threadIdx.x >> 5); item < n_items; item += gridDim.x * 4)
    {
        const uint2 it = __ldg(a.items_stream + item);
        const int py = (int)it.x / a.npx, px = (int)it.x - py * a.npx;
        const int x0 = px * KFB_PATCH_X + (lane & 3) * 4, y = py * KFB_PATCH_Y + (lane >> 2);
        if (x0 >= a.X || y >= a.Y) continue;
        const int z0 = (int)(it.y & 0xffffu), z1 = (int)(it.y >> 16);
        uint4 *vp = reinterpret_cast<uint4 *>(a.vol) + ((size_t)(z0 - a.z_store0) * a.Y + y) * (a.X >> 2) + (x0 >> 2);
        for (int z = z0; z <= z1; z += KFB_STREAM_DEPTH, vp += KFB_STREAM_DEPTH * plane4)
        {
            uint4 w[KFB_STREAM_DEPTH];
#pragma unroll
            for (int i = 0; i < KFB_STREAM_DEPTH; ++i)
                if (z + i <= z1) w[i] = __ldcs(vp + i * plane4);
#pragma unroll
            for (int i = 0; i < KFB_STREAM_DEPTH; ++i)
                if (z + i <= z1)
                {
                    update_free_quad<COUNT>(a, wt, vp + i * plane4, w[i], n_upd, n_st);
                    if (COUNT) ++n_ld;
                }
        }
    }
    if (COUNT)
    {
        if (n_upd) atomicAdd(a.counter, (unsigned long long)n_upd);
        if (n_ld) atomicAdd(a.counter + 2, (unsigned long long)n_ld);
        if (n_st) atomicAdd(a.counter + 3, (unsigned long long)n_st);
    }
}

// ---- general items: the reference's per-voxel arithmetic, software-pipelined over the planes ----------------------
// One plane of one thread: every load the plane needs -- the voxel quad (unconditionally: inside a thread's plane
// range nearly every quad is updated) and one 16-byte table entry per voxel -- is issued before the first use, so a
// plane costs ONE memory latency (the round-1 sweep paid thresholds -> exact depth -> voxels -> weight table in
// turn).  Tried and dropped (profiles/README.md, r02 integrate experiments): a two-stage software pipeline with the
// next plane's loads in flight, register-staged (96-128 registers: the lost occupancy cost more than the overlap
// gained, and ptxas put both stages' loads on one scoreboard) and cp.async-staged through shared memory (slower
// still).  This path lives on thread-level parallelism: 64 registers, 32 warps per SM.
struct GenStage
{
    float cz[4];   // vc.z per voxel (+inf: projects outside the image)
    float d2[4];   // |vc|^2 as the reference computes it (for the band)
    float4 tb[4];  // {hi_z, lo_z, depth, 1/lambda} of the pixel each voxel lands on
    uint4 word;    // the four voxels
};
struct GenConst
{
    unsigned long long vs2, sxy, szz, fxy, cxy, magic2;
    unsigned int last_pix;
    float rtrunc;
};
__device__ __forceinline__ void gen_issue(const IntegrateArgs &a, const GenConst &g, unsigned long long xy[4], unsigned long long zz[2],
                                          const uint4 *vp, GenStage &s)
{
    s.word = __ldcs(vp);
#pragma unroll
    for (int k = 0; k < 4; ++k) xy[k] = ffma2(g.vs2, g.sxy, xy[k]);
    zz[0] = ffma2(g.vs2, g.szz, zz[0]);
    zz[1] = ffma2(g.vs2, g.szz, zz[1]);
    unpack2(zz[0], s.cz[0], s.cz[1]);
    unpack2(zz[1], s.cz[2], s.cz[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        const float r = mufu_rcp(s.cz[k]);
        const unsigned long long q = fmul2(pack2(r, r), xy[k]);
        const unsigned long long m = fadd2(ffma2(q, g.fxy, g.cxy), g.magic2);
        float mu, mv, vxk, vyk;
        unpack2(m, mu, mv);
        unpack2(xy[k], vxk, vyk);
        s.d2[k] = dot3c(vxk, vyk, s.cz[k], vxk, vyk, s.cz[k]);
        const int ui = __float_as_int(mu) - KFB_MAGIC_I;
        const int vi = __float_as_int(mv) - KFB_MAGIC_I;
        const bool ok = ((unsigned)ui < (unsigned)a.w) & ((unsigned)vi < (unsigned)a.h);
        // an out-of-image voxel is classified as "behind everything": vc.z = +inf fails `<= hi_z` and passes
        // `> lo_z` for whatever (clamped) table entry it reads
        s.cz[k] = ok ? s.cz[k] : __int_as_float(0x7f800000);
        s.tb[k] = __ldg(a.tab4 + min((unsigned int)(vi * a.w + ui), g.last_pix));
    }
}
template <bool COUNT>
__device__ __forceinline__ void gen_process(const IntegrateArgs &a, const GenConst &g, const float4 *__restrict__ wt, const GenStage &s,
                                            uint4 *vp, int x0, int y, int z, unsigned int &n_upd, unsigned int &n_st)
{
    float t[4];
    const float4 *tb = s.tb;
    bool band = false;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        t[k] = s.cz[k] <= tb[k].x ? 1.0f : (s.cz[k] > tb[k].y ? KFB_SKIP : KFB_BAND);
        band |= (t[k] == KFB_BAND);
    }
    if (band)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k)
        {
            // exact sdf (tsdf_volume.cu:63-71): only here, between the two conservative thresholds
            const float nsdf = __fmaf_rn(tb[k].w, __fsqrt_rn(s.d2[k]), -tb[k].z);
            const float tb = nsdf <= a.trunc ? fminf(1.f, __fmul_rn(g.rtrunc, -nsdf)) : KFB_SKIP;
            t[k] = t[k] == KFB_BAND ? tb : t[k];
        }
    }
    if (!((t[0] != KFB_SKIP) | (t[1] != KFB_SKIP) | (t[2] != KFB_SKIP) | (t[3] != KFB_SKIP))) return;
    const uint4 wd = s.word;
    uint4 o = wd;
    const unsigned int w0 = wd.x >> 16, w1 = wd.y >> 16, w2 = wd.z >> 16, w3 = wd.w >> 16, mw = (unsigned)a.max_weight;
    if ((wd.x == wd.y) & (wd.x == wd.z) & (wd.x == wd.w) & (t[0] == t[1]) & (t[0] == t[2]) & (t[0] == t[3]) & (w0 <= mw))
    {
        // four equal words receive the same tsdf (free space in front of the band): one update
        o.x = o.y = o.z = o.w = update_word_e(wd.x, t[0], wt[w0]);
    }
    else if (max(max(w0, w1), max(w2, w3)) <= mw)
    {
        const float4 e0 = wt[w0], e1 = wt[w1], e2 = wt[w2], e3 = wt[w3];
        const unsigned int n0 = update_word_e(wd.x, t[0], e0), n1 = update_word_e(wd.y, t[1], e1);
        const unsigned int n2 = update_word_e(wd.z, t[2], e2), n3 = update_word_e(wd.w, t[3], e3);
        o.x = t[0] != KFB_SKIP ? n0 : wd.x;
        o.y = t[1] != KFB_SKIP ? n1 : wd.y;
        o.z = t[2] != KFB_SKIP ? n2 : wd.z;
        o.w = t[3] != KFB_SKIP ? n3 : wd.w;
    }
    else
    {
        if (t[0] != KFB_SKIP) o.x = update_generic(wd.x, t[0], a.max_weight);
        if (t[1] != KFB_SKIP) o.y = update_generic(wd.y, t[1], a.max_weight);
        if (t[2] != KFB_SKIP) o.z = update_generic(wd.z, t[2], a.max_weight);
        if (t[3] != KFB_SKIP) o.w = update_generic(wd.w, t[3], a.max_weight);
    }
    if (COUNT) n_upd += (t[0] != KFB_SKIP) + (t[1] != KFB_SKIP) + (t[2] != KFB_SKIP) + (t[3] != KFB_SKIP);
    if ((o.x != wd.x) | (o.y != wd.y) | (o.z != wd.z) | (o.w != wd.w))
    {
        __stcs(vp, o);
        if (COUNT) ++n_st;
        // a voxel that turns negative here (it was not before) activates the bricks around it
        if (((o.x & ~wd.x) | (o.y & ~wd.y) | (o.z & ~wd.z) | (o.w & ~wd.w)) & 0x8000u)
            mark_bricks(a.bricks, a.bdirty, a.bx, a.by, a.bz, a.bz0, x0, y, z);
    }
}

#ifndef KFB_GEN_MINB
#define KFB_GEN_MINB 8 // blocks of 4 warps per SM the register budget is sized for (64 registers)
#endif
template <bool COUNT, bool SMEM, int MINB>
__global__ void __launch_bounds__(128, MINB) integrate_general_kernel(const IntegrateArgs a)
{
    __shared__ float4 s_wt[SMEM ? KFB_WTAB_SMEM : 1];
    if (SMEM)
    {
        for (int i = threadIdx.x; i <= a.max_weight; i += blockDim.x) s_wt[i] = __ldg(a.wtab + i);
        __syncthreads();
    }
    const float4 *wt = SMEM ? s_wt : a.wtab;
    const int lane = threadIdx.x & 31;
    const unsigned int n_items = __ldg(a.plan_counts + 1);
    const size_t plane4 = ((size_t)a.X * a.Y) >> 2;
    GenConst g;
    {
        const float sz = a.pose.R.m[8];
        g.vs2 = pack2(a.vsx, a.vsx); g.sxy = pack2(a.pose.R.m[2], a.pose.R.m[5]); g.szz = pack2(sz, sz);
        g.fxy = pack2(a.fx, a.fy); g.cxy = pack2(a.cx, a.cy); g.magic2 = pack2(KFB_MAGIC_F, KFB_MAGIC_F);
        g.last_pix = (unsigned int)(a.w * a.h - 1);
        g.rtrunc = rcp_fdividef(a.trunc);
    }
    unsigned int n_upd = 0, n_ld = 0, n_st = 0;
    for (unsigned int item = blockIdx.x * 4 + (threadIdx.x >> 5); item < n_items; item += gridDim.x * 4)
    {
        const uint2 it = __ldg(a.items_general + item);
        const int py = (int)it.x / a.npx, px = (int)it.x - py * a.npx;
        const int x0 = px * KFB_PATCH_X + (lane & 3) * 4, y = py * KFB_PATCH_Y + (lane >> 2);
        if (x0 >= a.X || y >= a.Y) continue;
        const int z0 = (int)(it.y & 0xffffu), z1 = (int)(it.y >> 16);
        const int zstart = a.zb + ((z0 - a.zb) / a.zchunk) * a.zchunk; // first plane of the item's chunk
        unsigned long long xy[4], zz[2];
        {
            int zfrom = zstart;
            if (item < (unsigned int)a.gstate_cap)
            {
                const unsigned long long *st = a.gstates + (size_t)item * 192 + lane;
#pragma unroll
                for (int k = 0; k < 4; ++k) xy[k] = __ldcg(st + k * 32);
                zz[0] = __ldcg(st + 128);
                zz[1] = __ldcg(st + 160);
            }
            else
            {
                // more general items than state slots (a scene that is all surface): this item replays the
                // reference's running sum from the first plane by itself
                float vx[4], vy[4], vz[4];
                const float pyf = __fmul_rn((float)y, a.vsy), pz = __fmul_rn(0.f, a.vsz);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    const float3 r = rot3(a.pose.R, __fmul_rn((float)(x0 + k), a.vsx), pyf, pz);
                    vx[k] = __fadd_rn(r.x, a.pose.t[0]); vy[k] = __fadd_rn(r.y, a.pose.t[1]); vz[k] = __fadd_rn(r.z, a.pose.t[2]);
                }
                zfrom = 1;
                if (a.use_jump && a.zb - 1 >= a.jump_min)
                {
                    jump4(vx, a.vsx, a.pose.R.m[2], a.zb - 1);
                    jump4(vy, a.vsx, a.pose.R.m[5], a.zb - 1);
                    jump4(vz, a.vsx, a.pose.R.m[8], a.zb - 1);
                    zfrom = a.zb;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) xy[k] = pack2(vx[k], vy[k]);
                zz[0] = pack2(vz[0], vz[1]);
                zz[1] = pack2(vz[2], vz[3]);
            }
#pragma unroll 4
            for (int z = zfrom; z < z0; ++z)
            {
#pragma unroll
                for (int k = 0; k < 4; ++k) xy[k] = ffma2(g.vs2, g.sxy, xy[k]);
                zz[0] = ffma2(g.vs2, g.szz, zz[0]);
                zz[1] = ffma2(g.vs2, g.szz, zz[1]);
            }
        }
        uint4 *vp = reinterpret_cast<uint4 *>(a.vol) + ((size_t)(z0 - a.z_store0) * a.Y + y) * (a.X >> 2) + (x0 >> 2);
        // the pipelined path needs vc.z >= FLT_MIN on every visited plane (MUFU.RCP without the denormal
        // pre-scaling); vc.z is affine in z up to the running-sum drift, so the two ends decide with a 1 cm margin
        bool fast;
        {
            const float span = (float)(z1 - z0 + 1) * __fmul_rn(a.vsx, a.pose.R.m[8]);
            float c0, c1, c2, c3;
            unpack2(zz[0], c0, c1);
            unpack2(zz[1], c2, c3);
            const float m = fminf(fminf(c0, c1), fminf(c2, c3)); // plane z0 - 1
            fast = fminf(m, m + span) > 0.01f;
        }
        if (!fast)
        {
            // cold path (camera within a centimetre of this thread's planes): one plane at a time, any vc.z
            for (int z = z0; z <= z1; ++z, vp += plane4)
            {
#pragma unroll
                for (int k = 0; k < 4; ++k) xy[k] = ffma2(g.vs2, g.sxy, xy[k]);
                zz[0] = ffma2(g.vs2, g.szz, zz[0]);
                zz[1] = ffma2(g.vs2, g.szz, zz[1]);
                float cz[4], t[4];
                unpack2(zz[0], cz[0], cz[1]);
                unpack2(zz[1], cz[2], cz[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                {
                    float vxk, vyk;
                    unpack2(xy[k], vxk, vyk);
                    t[k] = classify_generic(a, vxk, vyk, cz[k], g.rtrunc);
                }
                if ((t[0] != KFB_SKIP) | (t[1] != KFB_SKIP) | (t[2] != KFB_SKIP) | (t[3] != KFB_SKIP))
                {
                    update_quad<COUNT>(a, vp, __ldcs(vp), t, x0, y, z, n_upd);
                    if (COUNT) { ++n_ld; ++n_st; }
                }
            }
            continue;
        }
        // ---- per-thread refinement of the item's plane range -------------------------------------------------
        // The plan decided for the whole 16 x 8 patch; a thread's four columns see a far smaller part of the image
        // (a few pixels), so its own free-space prefix [z0, fe_t] is longer and its own occlusion cut zb_t earlier
        // (same proofs as in the plan, with this thread's running sums after plane z0 - 1 as the affine base).
        // Each thread then walks its planes on its own: a cheap phase (tsdf = 1, no projection) and the full
        // per-voxel phase only on the planes in between -- for a surface seen at a grazing angle that is the
        // depth variation over the thread's few pixels instead of over the patch's footprint.
        int fe_t = z0 - 1, zb_t = z1;
        if (a.Sz > 1e-6f && a.refine)
        {
            float ax, ay, bx_, by_, c0, c1, c2, c3;
            unpack2(xy[0], ax, ay);
            unpack2(xy[3], bx_, by_);
            unpack2(zz[0], c0, c1);
            unpack2(zz[1], c2, c3);
            float umin = 1e30f, umax = -1e30f, vmin = 1e30f, vmax = -1e30f, zmin = 1e30f;
#pragma unroll
            for (int e = 0; e < 2; ++e)
            {
                const float zf = e ? (float)(z1 - z0 + 1) : 1.f; // planes z0 and z1, counted from the base plane z0 - 1
#pragma unroll
                for (int k = 0; k < 2; ++k)
                {
                    const float X = fmaf(zf, a.Sx, k ? bx_ : ax), Y = fmaf(zf, a.Sy, k ? by_ : ay), Zc = fmaf(zf, a.Sz, k ? c3 : c0);
                    const float r = mufu_rcp(fmaxf(Zc, 1e-3f));
                    const float u = fmaf(a.fx * X, r, a.cx), v = fmaf(a.fy * Y, r, a.cy);
                    umin = fminf(umin, u); umax = fmaxf(umax, u);
                    vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
                    zmin = fminf(zmin, Zc);
                }
            }
            if (zmin > 0.05f)
            {
                const float pad = 1.5f + (a.fx + a.fy + (float)(a.w + a.h)) * a.driftE * (1.001f * mufu_rcp(zmin));
                const bool all_inside = umin - pad >= 0.f && umax + pad <= (float)(a.w - 1) && vmin - pad >= 0.f && vmax + pad <= (float)(a.h - 1);
                const int u0 = max((int)floorf(fmaxf(umin - pad, -1e6f)), 0), u1 = min((int)ceilf(fminf(umax + pad, 1e6f)), a.w - 1);
                const int v0 = max((int)floorf(fmaxf(vmin - pad, -1e6f)), 0), v1 = min((int)ceilf(fminf(vmax + pad, 1e6f)), a.h - 1);
                if (u0 > u1 || v0 > v1) zb_t = z0 - 1; // never inside the image on these planes
                else
                {
                    const int span = max(u1 - u0, v1 - v0) + 1;
                    const int l = max(32 - __clz(span - 1) - 2, 2);
                    if (l <= 7)
                    {
                        const float2 *m = a.zmip + a.mip_off[l - 2];
                        const int mw = a.mip_w[l - 2];
                        const int tx0 = u0 >> l, tx1 = u1 >> l, ty1 = v1 >> l;
                        float2 q = make_float2(-1.f, 3.0e38f);
                        for (int ty = v0 >> l; ty <= ty1; ++ty)
                        {
                            float2 tl[5]; // a row of tiles in flight (the rectangle is at most 5 tiles wide)
#pragma unroll
                            for (int i = 0; i < 5; ++i) tl[i] = __ldg(m + ty * mw + min(tx0 + i, tx1));
#pragma unroll
                            for (int i = 0; i < 5; ++i) q = mm2(q, tl[i]);
                        }
                        const float zmin0 = fminf(c0, c3), zmax0 = fmaxf(c0, c3), e2 = 2.f * a.driftE;
                        const float zc = (q.x + e2 - zmin0) * a.invSz + 1.f;
                        zb_t = min(z1, z0 - 1 + (int)ceilf(fminf(zc, 1e6f)));
                        if (all_inside && !a.no_fastpath)
                        {
                            int zf = min(zb_t - (z0 - 1), (int)floorf(fminf(fmaxf((q.y - e2 - zmax0) * a.invSz, -1.f), 1e6f)));
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                if (zf >= 1 && fmaf((float)zf, a.Sz, zmax0) + e2 > q.y) --zf;
                            if (zf >= 1 && fmaf((float)zf, a.Sz, zmax0) + e2 <= q.y) fe_t = z0 - 1 + zf;
                        }
                    }
                }
            }
        }
        // ---- cheap phase: planes z0 .. fe_t are deep free space for this thread's four columns -----------------------
        int z = z0;
        for (; z <= fe_t; z += 4, vp += 4 * plane4)
        {
            uint4 w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (z + i <= fe_t) w[i] = __ldcs(vp + i * plane4);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (z + i <= fe_t)
                {
#pragma unroll
                    for (int k = 0; k < 4; ++k) xy[k] = ffma2(g.vs2, g.sxy, xy[k]);
                    zz[0] = ffma2(g.vs2, g.szz, zz[0]);
                    zz[1] = ffma2(g.vs2, g.szz, zz[1]);
                    update_free_quad<COUNT>(a, wt, vp + i * plane4, w[i], n_upd, n_st);
                    if (COUNT) ++n_ld;
                }
        }
        if (fe_t >= z0)
        {
            vp -= (size_t)(z - (fe_t + 1)) * plane4; // the loop stepped past fe_t in units of four planes
            z = fe_t + 1;
        }
        if (z > zb_t) continue;
        // ---- full phase: the exact per-voxel predicate, one plane (one memory latency) at a time ---------------------------
        for (; z <= zb_t; ++z, vp += plane4)
        {
            GenStage S;
            gen_issue(a, g, xy, zz, vp, S);
            gen_process<COUNT>(a, g, wt, S, vp, x0, y, z, n_upd, n_st);
            if (COUNT) ++n_ld;
        }
    }
    if (COUNT)
    {
        if (n_upd) atomicAdd(a.counter, (unsigned long long)n_upd);
        if (n_ld) atomicAdd(a.counter + 2, (unsigned long long)n_ld);
        if (n_st) atomicAdd(a.counter + 3, (unsigned long long)n_st);
    }
}

// Conservative frustum planes for this launch (see CullPlane).  vc(x, y, z) = P0(x, y) + z * S in exact
// arithmetic; the float running sum drifts from it by at most E = planes * 2^-24 * max|vc| per component.
static void make_cull_planes(const kfb_ctx *ctx, const IntegrateArgs &a, CullPlane out[KFB_NCULL])
{
    const Intr &k = ctx->L[0].k;
    const double S[3] = {(double)a.vsx * a.pose.R.m[2], (double)a.vsx * a.pose.R.m[5], (double)a.vsx * a.pose.R.m[8]};
    const double M = fabs(a.pose.t[0]) + fabs(a.pose.t[1]) + fabs(a.pose.t[2]) +
                     (double)ctx->p.volu_range[0] + ctx->p.volu_range[1] + ctx->p.volu_range[2] +
                     (double)a.vsx * ctx->p.volu_dims[2] * 1.01;
    const double E = ((double)ctx->p.volu_dims[2] + 16.0) * 1.2e-7 * M;
    const double mp = 0.01; // pixels: float error of the projection arithmetic
    const bool sane = k.cx >= 0.f && k.cx <= (float)(k.w - 1) && k.cy >= 0.f && k.cy <= (float)(k.h - 1) && k.fx > 0.f && k.fy > 0.f;
    const double pl[KFB_NCULL][3] = {
        {(double)k.fx, 0.0, (double)k.cx + 0.5 + mp},                // u >= -0.5
        {-(double)k.fx, 0.0, (double)k.w - 0.5 + mp - (double)k.cx}, // u <= w - 0.5
        {0.0, (double)k.fy, (double)k.cy + 0.5 + mp},                // v >= -0.5
        {0.0, -(double)k.fy, (double)k.h - 0.5 + mp - (double)k.cy}, // v <= h - 0.5
        {0.0, 0.0, 1.0},                                             // vc.z > 0
        {0.0, 0.0, -1.0},                                            // vc.z <= *zexit (added on the device)
    };
    for (int c = 0; c < KFB_NCULL; ++c)
    {
        CullPlane &o = out[c];
        const double n1 = fabs(pl[c][0]) + fabs(pl[c][1]) + fabs(pl[c][2]);
        const double g1 = pl[c][0] * S[0] + pl[c][1] * S[1] + pl[c][2] * S[2];
        o.a = (float)pl[c][0]; o.b = (float)pl[c][1]; o.g = (float)pl[c][2];
        o.slack = (float)((E + 2e-6 * M) * n1 * 1.01 + 1e-6);
        o.ninv = 0.f;
        if (!sane && c < 4) { o.kind = 3; continue; }
        if (fabs(g1) < 1e-9 * n1) o.kind = 2;
        else { o.kind = g1 > 0 ? 0 : 1; o.ninv = (float)(-1.0 / g1); }
    }
}

// Per-pixel tables of the current filtered depth (thresholds, exact operands, max-pyramid, zexit).  They depend
// on the depth image only, so the front end builds them on its own stream, off the frame's critical path.
int launch_build_tables(kfb_ctx *ctx, cudaStream_t stream)
{
    const Intr &k = ctx->L[0].k;
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->zexit, 0, sizeof(float), stream));
    dim3 b(32, 8), g((k.w + 31) / 32, (k.h + 7) / 8);
    build_tables_kernel<<<g, b, 0, stream>>>(ctx->L[0].depth, k.w, k.h, k.fx, k.fy, k.cx, k.cy, ctx->p.volu_trun_dist, ctx->tab_thrz,
                                            ctx->tab_exact, ctx->tab4, ctx->zexit);
    KFB_LAUNCH_CHECK(ctx);
    dim3 mg((k.w + 127) / 128, (k.h + 127) / 128);
    build_zmip_kernel<<<mg, 256, 0, stream>>>(ctx->tab_thrz, k.w, k.h, ctx->zmip, ctx->mip_off[0], ctx->mip_off[1], ctx->mip_off[2],
                                              ctx->mip_off[3], ctx->mip_off[4], ctx->mip_off[5]);
    KFB_LAUNCH_CHECK(ctx);
    {
        const size_t n0 = (size_t)k.w * k.h;
        dim3 sg((k.w + 31) / 32, (k.h + 7) / 8);
        for (int lv = 1; lv <= 6; ++lv)
        {
            build_zsparse_kernel<<<sg, 256, 0, stream>>>(ctx->tab_thrz, lv == 1 ? nullptr : ctx->zsparse + (size_t)(lv - 2) * n0,
                                                        ctx->zsparse + (size_t)(lv - 1) * n0, k.w, k.h, 1 << (lv - 1));
            KFB_LAUNCH_CHECK(ctx);
        }
    }
    return KFB_OK;
}

// round-1 sweep: one kernel that decides and sweeps per warp and z-chunk (kept behind KFB_INTEGRATE_V1=1 as the
// A/B partner of the planned sweep)
static int launch_integrate_v1(kfb_ctx *ctx, IntegrateArgs &a, int planes, uint64_t *n_updated)
{
    // z-chunks give resident warps and load balance (the visited interval differs per column); the running
    // sum at each chunk start comes from column_states_kernel, so chunks cost no replay.  48 B per thread per
    // chunk of state: bounded to max(128 MB, 1/8 of the volume).  KFB_INTEGRATE_ZCHUNKS overrides for tuning.
    const size_t nthr = (size_t)(a.X >> 2) * a.Y;
    int zc = (planes + 15) / 16;
    if (zc > 32) zc = 32;
    const size_t state_cap = std::max((size_t)128 << 20, ctx->vol_voxels * sizeof(uint32_t) / 8); // <= 1/8 of the volume (measured: 1024^3 wants 32 chunks, 403 MB)
    while (zc > 1 && (size_t)zc * 48 * nthr > state_cap) --zc;
    if (const char *e = getenv("KFB_INTEGRATE_ZCHUNKS")) zc = atoi(e) > 0 ? atoi(e) : zc;
    if (zc > planes) zc = planes;
    a.zchunk = (planes + zc - 1) / zc;
    a.nchunks = zc;
    {
        const size_t need = (size_t)zc * 48 * nthr + nthr * sizeof(int2);
        if (need > ctx->states_bytes)
        {
            KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->states) cudaFree(ctx->states);
            ctx->states = nullptr; ctx->states_bytes = 0;
            KFB_CUDA(ctx, cudaMalloc(&ctx->states, need));
            ctx->states_bytes = need;
        }
        a.states = ctx->states;
        a.col_range = reinterpret_cast<int2 *>(ctx->states + (size_t)zc * 6 * nthr);
        dim3 sb(32, 4), sg((a.X + 127) / 128, (a.Y + 3) / 4);
        column_states_kernel<<<sg, sb, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
    }
    dim3 block(32, KFB_INT_WARPS), grid((a.X + 4 * KFB_INT_WARPS * KFB_INT_PX - 1) / (4 * KFB_INT_WARPS * KFB_INT_PX), (a.Y + 32 / KFB_INT_PX - 1) / (32 / KFB_INT_PX), zc);
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
        // planes per iteration of the general path: with the stages' loads batched and short chunks, one plane
        // per iteration (fewest live registers) measured 2-3 % faster than two at 512^3, 1024^3 and 2048^3
        const int U = getenv("KFB_INTEGRATE_U") ? atoi(getenv("KFB_INTEGRATE_U")) : 1;
        if (U == 2) integrate_kernel<2, false><<<grid, block, 0, ctx->stream>>>(a);
        else if (U == 4) integrate_kernel<4, false><<<grid, block, 0, ctx->stream>>>(a);
        else integrate_kernel<1, false><<<grid, block, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        if (ctx->profiling) cudaEventRecord(ctx->events[61], ctx->stream);
    }
    return KFB_OK;
}

// planned sweep: plan -> states of the general items -> general items || stream items
static int launch_integrate_planned(kfb_ctx *ctx, IntegrateArgs &a, int planes, uint64_t *n_updated)
{
    a.refine = getenv("KFB_INTEGRATE_REFINE") ? 1 : 0;
    a.zchunk = KFB_PLAN_ZCHUNK;
    if (const char *e = getenv("KFB_PLAN_ZCHUNK")) { const int v = atoi(e); if (v >= 2 && v <= 64) a.zchunk = v; }
    a.nchunks = (planes + a.zchunk - 1) / a.zchunk;
    a.npx = (a.X + KFB_PATCH_X - 1) / KFB_PATCH_X;
    a.npy = (a.Y + KFB_PATCH_Y - 1) / KFB_PATCH_Y;
    a.mask_words = (a.nchunks + 31) / 32;
    if (a.ze > 65535) { ctx->err = "volumes deeper than 65535 planes are not supported"; return KFB_ERR_UNSUPPORTED; }
    const size_t npatch = (size_t)a.npx * a.npy, ncell = npatch * a.nchunks;
    const size_t gcap = std::max<size_t>(4096, ncell / 4);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_counts = 0, o_mask = 256, o_slot = o_mask + up(npatch * a.mask_words * 4), o_is = o_slot + up(ncell * 4),
                 o_ig = o_is + up(ncell * 8), o_st = o_ig + up(ncell * 8), need = o_st + gcap * 192 * 8;
    if (need > ctx->plan_bytes)
    {
        KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->plan_buf) cudaFree(ctx->plan_buf);
        ctx->plan_buf = nullptr; ctx->plan_bytes = 0;
        KFB_CUDA(ctx, cudaMalloc(&ctx->plan_buf, need));
        ctx->plan_bytes = need;
    }
    char *pb = (char *)ctx->plan_buf;
    a.plan_counts = (unsigned int *)(pb + o_counts);
    a.patch_mask = (unsigned int *)(pb + o_mask);
    a.slot_of = (unsigned int *)(pb + o_slot);
    a.items_stream = (uint2 *)(pb + o_is);
    a.items_general = (uint2 *)(pb + o_ig);
    a.gstates = (unsigned long long *)(pb + o_st);
    a.gstate_cap = (int)std::min<size_t>(gcap, 0x7fffffff);
    KFB_CUDA(ctx, cudaMemsetAsync(pb, 0, o_slot, ctx->stream)); // counters + masks
    if (n_updated) KFB_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, 4 * sizeof(unsigned long long), ctx->stream));
    integrate_plan_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    if (ctx->profiling) cudaEventRecord(ctx->events[60], ctx->stream);
    const bool smem = a.max_weight < KFB_WTAB_SMEM;
    // The general items are latency- and issue-bound and move few bytes, the stream items are bandwidth-bound: side
    // by side on two streams they fill each other's gaps (KFB_INTEGRATE_SERIAL=1: one after the other).  The stream
    // kernel is persistent with half an SM's warp slots; the general kernel gets about one block per four items
    // (sized from the previous frame's count; it strides, so any grid is correct) and takes whatever the SMs have
    // left, all of them once the stream items are done.
    const bool overlap = !getenv("KFB_INTEGRATE_SERIAL");
    int gs = ctx->sm_count * 8, gg = ctx->sm_count * KFB_GEN_MINB;
    if (!getenv("KFB_INTEGRATE_PERSISTENT"))
    {
        // about one warp per item, sized from the previous frame's counts (both kernels stride, so any grid is
        // correct): the block scheduler balances the two kernels over whatever the SMs have free
        const size_t hs = ctx->plan_hint_host ? (size_t)ctx->plan_hint_host[0] : 0, hg = ctx->plan_hint_host ? (size_t)ctx->plan_hint_host[1] : 0;
        const size_t cap = std::max<size_t>((ncell + 3) / 4, 1);
        gs = (int)std::min<size_t>(std::max<size_t>((hs + hs / 4 + 3) / 4, (size_t)gs), cap);
        gg = (int)std::min<size_t>(std::max<size_t>((hg + hg / 4 + 3) / 4, (size_t)gg), cap);
    }
    cudaStream_t gstr = ctx->stream;
    if (overlap)
    {
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_ifork, ctx->stream));
        KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->istream, ctx->ev_ifork, 0));
        gstr = ctx->istream;
    }
    // the running sums of the general items (FMA-pipe bound) run next to the stream items (bandwidth bound)
    integrate_states_kernel<<<(unsigned)((npatch + 3) / 4), 128, 0, gstr>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    if (n_updated)
    {
        if (smem) integrate_general_kernel<true, true, KFB_GEN_MINB><<<gg, 128, 0, gstr>>>(a);
        else integrate_general_kernel<true, false, KFB_GEN_MINB><<<gg, 128, 0, gstr>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        if (smem) integrate_stream_kernel<true, true><<<gs, 128, 0, ctx->stream>>>(a);
        else integrate_stream_kernel<true, false><<<gs, 128, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
    }
    else
    {
        const int minb = getenv("KFB_GEN_MINB") ? atoi(getenv("KFB_GEN_MINB")) : KFB_GEN_MINB; // register budget variant (tuning)
        if (!smem) integrate_general_kernel<false, false, KFB_GEN_MINB><<<gg, 128, 0, gstr>>>(a);
        else if (minb == 5) integrate_general_kernel<false, true, 5><<<gg, 128, 0, gstr>>>(a);
        else if (minb == 6) integrate_general_kernel<false, true, 6><<<gg, 128, 0, gstr>>>(a);
        else integrate_general_kernel<false, true, 8><<<gg, 128, 0, gstr>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        if (smem) integrate_stream_kernel<false, true><<<gs, 128, 0, ctx->stream>>>(a);
        else integrate_stream_kernel<false, false><<<gs, 128, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
    }
    if (overlap)
    {
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_ijoin, ctx->istream));
        KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ijoin, 0));
    }
    if (ctx->profiling) cudaEventRecord(ctx->events[61], ctx->stream);
    // item counts of this frame -> pinned host memory, read (one frame late, unsynchronised) as the next grid hint
    if (ctx->plan_hint_host)
        KFB_CUDA(ctx, cudaMemcpyAsync(ctx->plan_hint_host, a.plan_counts, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    if (n_updated)
    {
        KFB_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host, ctx->counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        KFB_CUDA(ctx, cudaMemcpyAsync(ctx->counters_host + 4, a.plan_counts, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *n_updated = ctx->counters_host[0];
    }
    return KFB_OK;
}

// everything about a sweep that does not depend on how it is cut into work
static void fill_integrate_args(kfb_ctx *ctx, const float vol2cam12[12], IntegrateArgs &a)
{
    const Intr &k = ctx->L[0].k;
    memset(&a, 0, sizeof(a));
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
    a.thrz = ctx->tab_thrz;
    a.exact = ctx->tab_exact;
    a.wtab = ctx->wtab;
    a.zexit = ctx->zexit;
    a.max_weight = ctx->p.tsdf_max_weight;
    a.no_fastpath = getenv("KFB_INTEGRATE_NOFAST") ? 1 : 0;
    a.no_prefix = getenv("KFB_INTEGRATE_NOPREFIX") ? 1 : 0;
    a.diag = getenv("KFB_INTEGRATE_DIAG") ? atoi(getenv("KFB_INTEGRATE_DIAG")) : 0;
    a.use_jump = getenv("KFB_INTEGRATE_NOJUMP") ? 0 : 1;
    // measured on B200: a jump costs about as much as 600 replayed planes (warps that straddle vc.x == 0 walk
    // many binades), so it pays for far z-slabs / large volumes, not for the z-chunks of a 512^3 sweep
    a.jump_min = getenv("KFB_INTEGRATE_JUMPMIN") ? atoi(getenv("KFB_INTEGRATE_JUMPMIN")) : 640;
    a.bricks = ctx->bricks;
    a.bdirty = ctx->bdirty;
    a.bx = ctx->bdim[0]; a.by = ctx->bdim[1]; a.bz = ctx->bdim[2]; a.bz0 = ctx->bz0;
    a.counter = ctx->counters;
    make_cull_planes(ctx, a, a.cull);
    a.zmip = ctx->zmip;
    for (int i = 0; i < 6; ++i) { a.mip_off[i] = ctx->mip_off[i]; a.mip_w[i] = (k.w + (1 << (i + 2)) - 1) >> (i + 2); }
    a.Sx = a.vsx * a.pose.R.m[2]; a.Sy = a.vsx * a.pose.R.m[5]; a.Sz = a.vsx * a.pose.R.m[8];
    a.invSz = a.Sz > 1e-6f ? 1.f / a.Sz : 0.f;
    {
        const double M = fabs(a.pose.t[0]) + fabs(a.pose.t[1]) + fabs(a.pose.t[2]) + (double)ctx->p.volu_range[0] + ctx->p.volu_range[1] +
                         ctx->p.volu_range[2] + (double)a.vsx * ctx->p.volu_dims[2] * 1.01;
        a.driftE = (float)(((double)ctx->p.volu_dims[2] + 16.0) * 1.2e-7 * M + 4e-6 * M); // as in make_cull_planes + float model evaluation
    }
    if (getenv("KFB_INTEGRATE_NOCULL") || getenv("KFB_INTEGRATE_NOOCC")) a.Sz = 0.f;
    if (getenv("KFB_INTEGRATE_NOCULL"))
        for (int c = 0; c < KFB_NCULL; ++c) a.cull[c].kind = 3;

    a.tab4 = ctx->tab4;
    a.zsparse = ctx->zsparse;
    a.use_sparse = (ctx->zsparse && !getenv("KFB_PLAN_TILES")) ? 1 : 0;
}

int launch_integrate(kfb_ctx *ctx, const float vol2cam12[12], uint64_t *n_updated)
{
    if (ctx->profiling) cudaEventRecord(ctx->events[56], ctx->stream); // whole call: 56 .. 57
    IntegrateArgs a;
    fill_integrate_args(ctx, vol2cam12, a);
    const int planes = a.ze - a.zb;
    if (planes <= 0) return KFB_OK;
    if (!getenv("KFB_INTEGRATE_V1"))
    {
        const int rcp = launch_integrate_planned(ctx, a, planes, n_updated);
        if (rcp) return rcp;
    }
    else
    {
        const int rc1 = launch_integrate_v1(ctx, a, planes, n_updated);
        if (rc1) return rc1;
    }
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_tables_free, ctx->stream)); // the next frame's tables may now be built
    const int rcd = launch_brick_distance(ctx);
    if (ctx->profiling) cudaEventRecord(ctx->events[57], ctx->stream);
    return rcd;
}

int launch_build_wtab(kfb_ctx *ctx)
{
    const int n = ctx->p.tsdf_max_weight + 1;
    build_wtab_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->wtab, ctx->p.tsdf_max_weight);
    KFB_LAUNCH_CHECK(ctx);
    return KFB_OK;
}

// ---- brick map maintenance -------------------------------------------------------------------------
// full rescan (after kfb_upload_volume): same marking rule as the integrate kernel
__global__ void rebuild_bricks_kernel(const IntegrateArgs a, int zs0, int zs1)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= a.X || y >= a.Y) return;
    const size_t plane4 = ((size_t)a.X * a.Y) >> 2;
    const uint4 *vp = reinterpret_cast<const uint4 *>(a.vol) + (size_t)y * (a.X >> 2) + (x0 >> 2);
    for (int z = zs0 + blockIdx.z; z < zs1; z += gridDim.z)
    {
        const uint4 o = __ldg(vp + (size_t)(z - zs0) * plane4);
        if ((o.x | o.y | o.z | o.w) & 0x8000u) mark_bricks(a.bricks, a.bdirty, a.bx, a.by, a.bz, a.bz0, x0, y, z);
    }
}

int launch_rebuild_bricks(kfb_ctx *ctx)
{
    const size_t nb = (size_t)ctx->bdim[0] * ctx->bdim[1] * ctx->bdim[2];
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->bricks, 0, nb, ctx->stream));
    IntegrateArgs a;
    memset(&a, 0, sizeof(a));
    a.vol = ctx->vol;
    a.X = ctx->p.volu_dims[0]; a.Y = ctx->p.volu_dims[1];
    a.bricks = ctx->bricks;
    a.bdirty = ctx->bdirty;
    a.bx = ctx->bdim[0]; a.by = ctx->bdim[1]; a.bz = ctx->bdim[2]; a.bz0 = ctx->bz0;
    dim3 block(32, 4), grid((a.X / 4 + 31) / 32, (a.Y + 3) / 4, 32);
    rebuild_bricks_kernel<<<grid, block, 0, ctx->stream>>>(a, ctx->z0, ctx->z1);
    KFB_LAUNCH_CHECK(ctx);
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->bdirty, 1, sizeof(int), ctx->stream)); // non-zero: force a distance rebuild
    return launch_brick_distance(ctx);
}

// ---- brick distance map ----------------------------------------------------------------------------
// bdist[b] = min(KFB_BDIST_CAP, Chebyshev distance in bricks from b to the nearest active brick), 0 for an
// active brick.  Three separable passes: d(b) = min over offsets j along the axis of max(src(b + j), |j|).
// Every pass returns at once unless a brick flag has flipped since the last rebuild (*dirty != 0); flags
// only ever flip 0 -> 1 between resets, so a clean map stays exact.
__global__ void brick_distance_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int bx, int by, int bz, int axis,
                                      const int *__restrict__ dirty, int from_flags)
{
    if (*dirty == 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = bx * by * bz;
    if (i >= n) return;
    const int x = i % bx, y = (i / bx) % by, z = i / (bx * by);
    const int pos = axis == 0 ? x : (axis == 1 ? y : z);
    const int len = axis == 0 ? bx : (axis == 1 ? by : bz);
    const int stride = axis == 0 ? 1 : (axis == 1 ? bx : bx * by);
    int best = KFB_BDIST_CAP;
    const int j0 = max(-KFB_BDIST_CAP, -pos), j1 = min(KFB_BDIST_CAP, len - 1 - pos);
    for (int j = j0; j <= j1; ++j)
    {
        int v = src[i + j * stride];
        if (from_flags) v = v ? 0 : KFB_BDIST_CAP;
        best = min(best, max(v, abs(j)));
    }
    dst[i] = (uint8_t)best;
}

int launch_brick_distance(kfb_ctx *ctx)
{
    const int n = ctx->bdim[0] * ctx->bdim[1] * ctx->bdim[2];
    const int blocks = (n + 255) / 256;
    brick_distance_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bricks, ctx->bdist_tmp, ctx->bdim[0], ctx->bdim[1], ctx->bdim[2], 0, ctx->bdirty, 1);
    KFB_LAUNCH_CHECK(ctx);
    brick_distance_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bdist_tmp, ctx->bdist_tmp2, ctx->bdim[0], ctx->bdim[1], ctx->bdim[2], 1, ctx->bdirty, 0);
    KFB_LAUNCH_CHECK(ctx);
    brick_distance_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bdist_tmp2, ctx->bdist, ctx->bdim[0], ctx->bdim[1], ctx->bdim[2], 2, ctx->bdirty, 0);
    KFB_LAUNCH_CHECK(ctx);
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->bdirty, 0, sizeof(int), ctx->stream));
    return KFB_OK;
}

// ---- work histogram over planes (slab balancing for sharded volumes) -----------------------------------------
// Number of 4-voxel groups the sweep would visit on every plane of the WHOLE volume for this depth image and
// pose (frustum interval only; independent of the volume's content and of the planes this context stores).
// Written as a difference array: +1 at the first visited plane, -1 after the last.
__global__ void __launch_bounds__(128) plane_histogram_kernel(const IntegrateArgs a, int Z, int *__restrict__ diff)
{
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 4 + threadIdx.y;
    if (x0 >= a.X || y >= a.Y) return;
    float vx[2], vy[2], vz[2];
    const float py = __fmul_rn((float)y, a.vsy);
    const float pz = __fmul_rn(0.f, a.vsz);
#pragma unroll
    for (int k = 0; k < 2; ++k)
    {
        const float px = __fmul_rn((float)(x0 + 3 * k), a.vsx);
        const float3 r = rot3(a.pose.R, px, py, pz);
        vx[k] = __fadd_rn(r.x, a.pose.t[0]);
        vy[k] = __fadd_rn(r.y, a.pose.t[1]);
        vz[k] = __fadd_rn(r.z, a.pose.t[2]);
    }
    float lo, hi;
    frustum_interval(a, vx[0], vy[0], vz[0], vx[1], vy[1], vz[1], 1, Z, lo, hi);
    const int za = max(1, (int)floorf(lo)), zb = min(Z - 1, (int)ceilf(hi));
    if (za > zb) return;
    atomicAdd(diff + za, 1);
    atomicAdd(diff + zb + 1, -1);
}

static int launch_plane_histogram_frustum(kfb_ctx *ctx, const float vol2cam12[12], uint32_t *host_hist)
{
    const Intr &k = ctx->L[0].k;
    const int Z = ctx->p.volu_dims[2];
    IntegrateArgs a;
    memset(&a, 0, sizeof(a));
    a.X = ctx->p.volu_dims[0]; a.Y = ctx->p.volu_dims[1];
    a.pose = make_pose(vol2cam12);
    a.vsx = ctx->voxel_size[0]; a.vsy = ctx->voxel_size[1]; a.vsz = ctx->voxel_size[2];
    a.fx = k.fx; a.fy = k.fy; a.cx = k.cx; a.cy = k.cy; a.w = k.w; a.h = k.h;
    a.zexit = ctx->zexit;
    make_cull_planes(ctx, a, a.cull);
    int *diff = nullptr;
    KFB_CUDA(ctx, cudaMalloc(&diff, (size_t)(Z + 2) * sizeof(int)));
    KFB_CUDA(ctx, cudaMemsetAsync(diff, 0, (size_t)(Z + 2) * sizeof(int), ctx->stream));
    dim3 block(32, 4), grid((a.X + 127) / 128, (a.Y + 3) / 4);
    plane_histogram_kernel<<<grid, block, 0, ctx->stream>>>(a, Z, diff);
    ctx->launches++;
    std::vector<int> h((size_t)Z + 2);
    cudaError_t e = cudaMemcpyAsync(h.data(), diff, h.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(diff);
    KFB_CUDA(ctx, e);
    long run = 0;
    for (int z = 0; z < Z; ++z) { run += h[z]; host_hist[z] = (uint32_t)(run > 0 ? run : 0); }
    return KFB_OK;
}

// Work per plane from the sweep's own plan: the plan kernel runs for the WHOLE volume (it needs the depth tables and
// the pose, not the voxels), and every item adds its 32 voxel quads per plane to the planes it covers -- once for a
// stream item, KFB_HIST_GENERAL_WEIGHT times for a general item (measured cost ratio per quad at 512^3).  The
// frustum-only count above left the busiest of two 2048^3 slabs 30 % behind the other (profiles/README.md).
#define KFB_HIST_GENERAL_WEIGHT 4
int launch_plane_histogram(kfb_ctx *ctx, const float vol2cam12[12], uint32_t *host_hist)
{
    const int Z = ctx->p.volu_dims[2];
    if (getenv("KFB_HIST_FRUSTUM")) return launch_plane_histogram_frustum(ctx, vol2cam12, host_hist);
    IntegrateArgs a;
    fill_integrate_args(ctx, vol2cam12, a);
    a.vol = nullptr;
    a.z_store0 = 0; a.zb = 1; a.ze = Z;
    if (Z > 65535) { ctx->err = "volumes deeper than 65535 planes are not supported"; return KFB_ERR_UNSUPPORTED; }
    a.zchunk = 16;
    a.nchunks = (Z - 1 + a.zchunk - 1) / a.zchunk;
    a.npx = (a.X + KFB_PATCH_X - 1) / KFB_PATCH_X;
    a.npy = (a.Y + KFB_PATCH_Y - 1) / KFB_PATCH_Y;
    a.mask_words = (a.nchunks + 31) / 32;
    const size_t npatch = (size_t)a.npx * a.npy, ncell = npatch * a.nchunks;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_mask = 256, o_slot = o_mask + up(npatch * a.mask_words * 4), o_is = o_slot + up(ncell * 4), o_ig = o_is + up(ncell * 8),
                 need = o_ig + up(ncell * 8);
    char *pb = nullptr;
    KFB_CUDA(ctx, cudaMalloc(&pb, need));
    a.plan_counts = (unsigned int *)pb;
    a.patch_mask = (unsigned int *)(pb + o_mask);
    a.slot_of = (unsigned int *)(pb + o_slot);
    a.items_stream = (uint2 *)(pb + o_is);
    a.items_general = (uint2 *)(pb + o_ig);
    cudaError_t e = cudaMemsetAsync(pb, 0, o_slot, ctx->stream);
    if (e == cudaSuccess)
    {
        integrate_plan_kernel<<<(unsigned)((ncell + 127) / 128), 128, 0, ctx->stream>>>(a);
        ctx->launches++;
        e = cudaGetLastError();
    }
    unsigned int counts[2] = {0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, a.plan_counts, sizeof(counts), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    std::vector<uint2> items[2];
    for (int t = 0; t < 2 && e == cudaSuccess; ++t)
    {
        items[t].resize(counts[t]);
        if (counts[t]) e = cudaMemcpy(items[t].data(), t ? a.items_general : a.items_stream, (size_t)counts[t] * sizeof(uint2), cudaMemcpyDeviceToHost);
    }
    cudaFree(pb);
    KFB_CUDA(ctx, e);
    std::vector<long long> diff((size_t)Z + 2, 0);
    for (int t = 0; t < 2; ++t)
    {
        const long long wgt = 32 * (t ? KFB_HIST_GENERAL_WEIGHT : 1);
        for (const uint2 &it : items[t])
        {
            const int z0 = (int)(it.y & 0xffffu), z1 = (int)(it.y >> 16);
            diff[z0] += wgt;
            diff[z1 + 1] -= wgt;
        }
    }
    long long run = 0;
    for (int z = 0; z < Z; ++z)
    {
        run += diff[z];
        host_hist[z] = (uint32_t)std::min<long long>(std::max<long long>(run, 0), 0xffffffffll);
    }
    return KFB_OK;
}

// resetVolume: zero every stored voxel (the reference's fixed 32x32 grid is a bug)
int launch_reset_volume(kfb_ctx *ctx)
{
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->vol, 0, ctx->vol_voxels * sizeof(uint32_t), ctx->stream));
    const size_t nb = (size_t)ctx->bdim[0] * ctx->bdim[1] * ctx->bdim[2];
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->bricks, 0, nb, ctx->stream));
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->bdist, KFB_BDIST_CAP, nb, ctx->stream)); // no active brick anywhere
    KFB_CUDA(ctx, cudaMemsetAsync(ctx->bdirty, 0, sizeof(int), ctx->stream));
    return KFB_OK;
}

} // namespace kfb
