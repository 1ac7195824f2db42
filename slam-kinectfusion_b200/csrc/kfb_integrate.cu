// kfb_integrate.cu -- TSDF integration (replaces kf::device::integrate,
// kfusion/src/tsdf_volume.cu:34-111) and volume reset (tsdf_volume.cu:11-32).
//
// A planned sweep over the packed {int16 tsdf, int16 weight} volume (8 x 8 x 8 voxel bricks, kfb_common.cuh: vol_index).
// A thread owns four consecutive x voxels (one 128-bit load and store per plane, issued only where the reference's
// update predicate can pass), a warp a 16 x 8 voxel patch (two bricks wide) that it marches along z.  The predicate and
// the update arithmetic are the reference's, rounding for rounding:
//   * vc(z) is the reference's running sum, vc = fma(voxel_size.x, R[:,2], vc), from z = 1, carried as packed f32x2
//     pairs (FFMA2);
//   * pixel = round-half-even(fma(vc.x * MUFU.RCP(vc.z), fx, cx)), done with the 2^23 magic-number add so it stays off
//     the conversion pipe;
//   * sdf = depth - |vc| / lambda is only evaluated exactly inside a narrow band around the surface: a per-pixel
//     table of conservative vc.z thresholds classifies "certainly free space => tsdf == 1.0f exactly" and "certainly
//     behind the surface => rejected" with two compares.  The thresholds are proved conservative in
//     build_tables_kernel, so classification never changes a result, only skips work;
//   * the update runs without the quarter-rate XU pipe (int->float by magic-number add, the reciprocal of weight + 1
//     from a device-built table of MUFU.RCP results, float->int truncation by an RZ add of 2^23); a quad whose four
//     words and tsdf values agree is updated once; stores whose value equals the loaded one are dropped.
// What is decided where:
//   * integrate_plan_kernel, one thread per (patch, 16-plane chunk): conservative frustum interval, occlusion cut and
//     deep-free-space prefix from an EXACT min/max of the thresholds over the pixel rectangle the patch projects into
//     (sparse table over 2^k windows) -> two item lists;
//   * integrate_stream_kernel: items whose every voxel gets tsdf = 1.0f -- load, running mean, store, a brick layer in
//     flight per thread (bandwidth-bound: 0.94 of the measured copy bandwidth when every voxel is updated);
//   * integrate_general_kernel: the exact per-voxel predicate, one plane per iteration, all of a plane's loads
//     requested before the first use, the item's planes prefetched into L2 when it starts; its first blocks walk the
//     running sums of all patches once and publish them per item (release / acquire), a far z-slab's prefix is
//     crossed by an exact integer jump (jump_fma);
//   * the two kernels run side by side on two streams (the general one with priority).
// Side product for the raycaster: a voxel that turns negative marks the 8^3 bricks within two voxels of it in a byte
// map; three separable passes turn the map into the brick distance field kfb_raycast.cu skips with.
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
    uint32_t *vol;  // brick-major packed voxels (kfb_common.cuh: vol_index)
    int X, Y;
    int zb, ze;     // global planes to process: [zb, ze), zb >= 1
    int zchunk;     // planes per plan chunk
    int nchunks;
    Pose pose;      // vol2cam
    float vsx, vsy, vsz;
    float trunc;
    float fx, fy, cx, cy;
    int w, h;
    const float4 *wtab;
    const float *zexit;
    float Sx, Sy, Sz, invSz, driftE; // per-plane step of vc (float), 1/Sz, bound on the running-sum drift
    int max_weight;
    int no_fastpath;            // KFB_INTEGRATE_NOFAST: no stream items (tuning / testing)
    int use_jump, jump_min;     // exact jump of the running sum for prefixes of at least jump_min planes
    uint8_t *bricks;
    int *bdirty;         // receives dirty_tag when a brick flag flips 0 -> 1 (the distance map must be rebuilt)
    int dirty_tag;       // this call's tag (never 0, never reused): nothing has to be cleared between frames
    int bx, by, bz, bz0; // brick grid dims (x, y, stored z bricks) and first stored z brick -- of the flags AND of the voxels
    CullPlane cull[KFB_NCULL];
    unsigned long long *counter;
    // ---- work plan (integrate_plan_kernel): the sweep as two lists of (patch, plane range) items ----
    const float4 *tab4;           // per pixel {hi_z, lo_z, depth, 1/lambda}
    const float2 *zsparse;        // sparse table of {max lo_z, min hi_z}: level k (1..6) at pixel (x, y) covers the 2^k x 2^k window from (x, y)
    int npx, npy;                 // 16 x 8 voxel patches in x and y
    int mask_words;               // 32-bit words of a patch's "chunk has a general item" mask
    uint2 *items_stream;          // {patch, z0 | z1 << 16}: every voxel of these planes gets tsdf = 1.0f
    uint2 *items_general;         // {patch, z0 | z1 << 16}: planes that need the per-voxel predicate (within one chunk)
    unsigned int *plan_counts;    // [0] stream items, [1] general items
    unsigned int *patch_mask;     // [patch][mask_words]
    unsigned int *slot_of;        // [patch][chunk] -> general item index
    unsigned long long *gstates;  // [general item][6][32]: packed vc of the item's 32 threads after plane zstart(chunk) - 1
    int gstate_cap;               // general items that have a state slot (the others replay their running sums)
    // fused mode: the first producer_blocks blocks of the general kernel walk the running sums and publish each item's
    // state with ready[item] = ready_tag (release); the item's warp waits for it (acquire)
    unsigned int *ready;
    unsigned int ready_tag;
    int producer_blocks;
    int gen_prefetch;    // general items ask L2 for all their planes at once, before the one-plane-at-a-time march
    unsigned long long *err;      // mapped host word: set when a wait gave up
};

#define KFB_MAGIC_F 12582912.0f   // 1.5 * 2^23
#define KFB_MAGIC_I 0x4B400000
#define KFB_SKIP (-4.0f)


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
                                    float cy, float trunc, float2 *__restrict__ thrz, float4 *__restrict__ tab4, float *__restrict__ zexit)
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

__device__ __forceinline__ float2 mm2(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fminf(a.y, b.y)); } // {max, min}
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
            const float4 e = __ldg(a.tab4 + p); // {hi_z, lo_z, depth, 1/lambda}
            if (e.z > 0.f)
            {
                const float d2 = dot3c(cx_, cy_, cz, cx_, cy_, cz);
                const float nsdf = __fmaf_rn(e.w, __fsqrt_rn(d2), -e.z);
                if (nsdf <= a.trunc) t = fminf(1.f, __fmul_rn(rtrunc, -nsdf));
            }
        }
    }
    return t;
}

// bricks within two voxels of a sample that just turned negative (see kfb_raycast.cu for why two)
__device__ __noinline__ void mark_bricks(uint8_t *flags, int *dirty, int dirty_tag, int gbx, int gby, int gbz, int gbz0, int x0, int y, int z)
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
    if (any) *dirty = dirty_tag;
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
            mark_bricks(a.bricks, a.bdirty, a.dirty_tag, a.bx, a.by, a.bz, a.bz0, x0, y, z);
    }
}

#define KFB_BAND (-8.0f)
// ---- the sweep as a work plan ---------------------------------------------------------------------------------
// The decisions a sweep needs -- which planes of which voxel columns can pass the reference's predicate at all, which
// of them are deep free space, which need the per-voxel predicate -- are taken once, by a small kernel with one thread
// per (16 x 8 voxel patch, 16-plane chunk), which writes two compact work lists (round 1 took them inside the sweep,
// per warp and chunk: with every warp slot taken by warps that wait on memory, set-up, streaming and per-voxel work
// simply added up):
//   stream items  {patch, z0..z1}: every voxel of these planes passes the predicate with tsdf == 1.0f exactly --
//                 integrate_stream_kernel does nothing but load -> running mean -> store, eight planes in flight;
//   general items {patch, z0..z1}: planes on which the exact per-voxel predicate decides (the band around the
//                 surface, the frustum border, holes) -- integrate_general_kernel runs the reference's arithmetic,
//                 one plane per iteration with all of the plane's loads issued together.
// The running sums of the reference (vc += zstep) are only needed by the general items: integrate_states_kernel
// walks the patches that have any, one warp per patch, and stores the 32 threads' sums at the chunk starts the
// plan asked for (48 B per thread per general item).  States and general items run on a second stream next to the
// stream items (FMA-bound, latency-bound and bandwidth-bound work side by side).
// Culling is conservative and the per-voxel predicate is exact, so no split of the work can change a result:
// tests/test_ref_ab.py and tests/test_ref_full.py compare whole volumes with the reference kernels' bit for bit.
//
// Voxel layout: 8 x 8 x 8 bricks (kfb_common.cuh: vol_index).  A warp's 16 x 8 patch is two bricks wide, so one plane
// of it is two contiguous 256-byte runs (lane = brick half * 16 + row * 2 + quad), and its z-march stays inside the
// same two bricks for eight planes.
#define KFB_PATCH_X 16
#define KFB_PATCH_Y 8
#define KFB_PLAN_ZCHUNK 16 // planes per plan chunk: two brick layers (measured at 512^3, brick-major: 6 / 8 / 16 planes -> sweep 122 / 112 / 104 us)

// A sweep thread of a patch: lane = brick half * 16 + row * 2 + quad.  x0 / y = its four voxels; quad_index(z) = index,
// in 16-byte units, of those voxels on plane z in the brick-major volume.
struct PatchLane
{
    int x0, y;
    size_t brick_xy; // index of the thread's brick column among the bricks of one brick layer
    int in_plane;    // (y & 7) * 2 + quad: offset inside a brick's plane, 16-byte units
};
__device__ __forceinline__ PatchLane patch_lane(const IntegrateArgs &a, int patch, int lane)
{
    const int py = patch / a.npx, px = patch - py * a.npx;
    const int half = lane >> 4, row = (lane >> 1) & 7, quad = lane & 1;
    PatchLane p;
    p.x0 = px * KFB_PATCH_X + half * 8 + quad * 4;
    p.y = py * KFB_PATCH_Y + row;
    p.brick_xy = (size_t)py * a.bx + (size_t)(2 * px + half);
    p.in_plane = row * 2 + quad;
    return p;
}
// plan chunks sit at absolute multiples of the chunk height (a multiple of the brick height by default), so that a
// chunk is made of whole brick layers; the first chunk starts at the first processed plane
__device__ __forceinline__ int chunk_first_plane(const IntegrateArgs &a, int c) { return max(a.zb, (a.zb / a.zchunk + c) * a.zchunk); }
__device__ __forceinline__ int chunk_end_plane(const IntegrateArgs &a, int c) { return min(a.ze, (a.zb / a.zchunk + c + 1) * a.zchunk); }
__device__ __forceinline__ size_t quad_index(const IntegrateArgs &a, const PatchLane &p, int z)
{
    return ((((size_t)((z >> 3) - a.bz0)) * ((size_t)a.bx * a.by) + p.brick_xy) << 7) + (size_t)(((z & 7) << 4) | p.in_plane);
}

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
        const int zstart = chunk_first_plane(a, c), zend = chunk_end_plane(a, c);
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
                    {
                        // exact cover of the rectangle by overlapping 2^k x 2^k windows of the sparse table: k from the
                        // shorter side, raised until five windows reach along the longer one
                        float2 q = make_float2(-1.f, 3.0e38f);
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
template <bool PUBLISH>
__device__ __forceinline__ void states_walk(const IntegrateArgs &a, int patch, int lane)
{
    if (patch >= a.npx * a.npy) return;
    int c_last = -1;
    for (int w = a.mask_words - 1; w >= 0 && c_last < 0; --w)
    {
        const unsigned m = __ldg(a.patch_mask + (size_t)patch * a.mask_words + w);
        if (m) c_last = w * 32 + 31 - __clz(m);
    }
    if (c_last < 0) return;
    // lanes beyond the volume's edge compute a valid neighbour's sums (never read)
    const PatchLane pl = patch_lane(a, patch, lane);
    const int x0 = min(pl.x0, a.X - 4), y = min(pl.y, a.Y - 1);
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
        const int target = chunk_first_plane(a, c) - 1; // state after this plane
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
        if (PUBLISH)
        {
            // the 32 lanes' stores, then one release store of the frame's tag: the item's warp may start
            __syncwarp();
            if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.ready + slot), "r"(a.ready_tag) : "memory");
        }
    }
}
__global__ void __launch_bounds__(128) integrate_states_kernel(const IntegrateArgs a)
{
    states_walk<false>(a, blockIdx.x * 4 + (threadIdx.x >> 5), threadIdx.x & 31);
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
    unsigned int n_upd = 0, n_ld = 0, n_st = 0;
    const unsigned int wpb = blockDim.x >> 5; // one item per warp
    for (unsigned int item = blockIdx.x * wpb + (threadIdx.x >> 5); item < n_items; item += gridDim.x * wpb)
    {
        const uint2 it = __ldg(a.items_stream + item);
        const PatchLane pl = patch_lane(a, (int)it.x, lane);
        if (pl.x0 >= a.X || pl.y >= a.Y) continue;
        const int z0 = (int)(it.y & 0xffffu), z1 = (int)(it.y >> 16);
        uint4 *const vol4 = reinterpret_cast<uint4 *>(a.vol);
        for (int zl = z0 & ~7; zl <= z1; zl += 8) // one brick layer at a time: its eight planes are 256 B apart
        {
            uint4 *const base = vol4 + quad_index(a, pl, zl);
            uint4 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (zl + i >= z0 && zl + i <= z1) w[i] = __ldcs(base + i * 16);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (zl + i >= z0 && zl + i <= z1)
                {
                    update_free_quad<COUNT>(a, wt, base + i * 16, w[i], n_upd, n_st);
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
            mark_bricks(a.bricks, a.bdirty, a.dirty_tag, a.bx, a.by, a.bz, a.bz0, x0, y, z);
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
    GenConst g;
    {
        const float sz = a.pose.R.m[8];
        g.vs2 = pack2(a.vsx, a.vsx); g.sxy = pack2(a.pose.R.m[2], a.pose.R.m[5]); g.szz = pack2(sz, sz);
        g.fxy = pack2(a.fx, a.fy); g.cxy = pack2(a.cx, a.cy); g.magic2 = pack2(KFB_MAGIC_F, KFB_MAGIC_F);
        g.last_pix = (unsigned int)(a.w * a.h - 1);
        g.rtrunc = rcp_fdividef(a.trunc);
    }
    unsigned int n_upd = 0, n_ld = 0, n_st = 0;
    const unsigned int wpb = blockDim.x >> 5; // one item per warp
    // Fused mode: the first producer_blocks blocks walk the running sums (one warp per patch, chunk by chunk) while the
    // blocks behind them already work on the items whose states are there -- the walk (22 us, FMA-bound) no longer
    // precedes the items, it runs under them.  Blocks are dispatched in index order, so every producer is resident
    // (or waiting for a slot that the independent stream kernel will free) before the first item block can spin.
    if ((int)blockIdx.x < a.producer_blocks)
    {
        states_walk<true>(a, (int)(blockIdx.x * wpb + (threadIdx.x >> 5)), lane);
        return;
    }
    const unsigned int cblock = blockIdx.x - (unsigned)a.producer_blocks, cgrid = gridDim.x - (unsigned)a.producer_blocks;
    for (unsigned int item = cblock * wpb + (threadIdx.x >> 5); item < n_items; item += cgrid * wpb)
    {
        const uint2 it = __ldg(a.items_general + item);
        const PatchLane pl = patch_lane(a, (int)it.x, lane);
        const int x0 = pl.x0, y = pl.y;
        if (x0 >= a.X || y >= a.Y) continue;
        const int z0 = (int)(it.y & 0xffffu), z1 = (int)(it.y >> 16);
        const int zstart = max(a.zb, (z0 / a.zchunk) * a.zchunk); // first plane of the item's chunk
        uint4 *const vol4 = reinterpret_cast<uint4 *>(a.vol);
        if (a.gen_prefetch)
        {
            // The march below pays one memory latency per plane.  All of the item's voxels are known now: per brick
            // layer the patch's two bricks are 32 lines of 128 bytes (brick half, plane, half plane = the lane's own
            // half, row, quad fields), so one prefetch per lane and layer brings every plane the item visits into L2
            // while the states are fetched and the first planes are processed.
            for (int zl = z0 & ~7; zl <= z1; zl += 8)
            {
                const int pz = zl + ((lane >> 1) & 7);
                if (pz >= z0 && pz <= z1)
                {
                    const uint4 *line = vol4 + ((((size_t)((zl >> 3) - a.bz0)) * ((size_t)a.bx * a.by) + pl.brick_xy) << 7) +
                                        (size_t)((((lane >> 1) & 7) << 4) | ((lane & 1) << 3));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                }
            }
        }
        unsigned long long xy[4], zz[2];
        {
            int zfrom = zstart;
            if (item < (unsigned int)a.gstate_cap)
            {
                if (a.producer_blocks > 0)
                {
                    if (lane == 0)
                    {
                        unsigned int v, spins = 0;
                        for (;;)
                        {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.ready + item) : "memory");
                            if (v == a.ready_tag) break;
                            __nanosleep(100);
                            if (++spins > (1u << 22)) { *(volatile unsigned long long *)a.err = 0x1000000000000000ull | item; break; } // ~0.5 s: report, do not hang
                        }
                    }
                    __syncwarp();
                }
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
        // the fast path needs vc.z >= FLT_MIN on every visited plane (MUFU.RCP without the denormal
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
            for (int z = z0; z <= z1; ++z)
            {
                uint4 *vp = vol4 + quad_index(a, pl, z);
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
        // the exact per-voxel predicate, one plane (one memory latency) at a time
        uint4 *vp = vol4 + quad_index(a, pl, z0);
        for (int z = z0; z <= z1; ++z)
        {
            GenStage S;
            gen_issue(a, g, xy, zz, vp, S);
            gen_process<COUNT>(a, g, wt, S, vp, x0, y, z, n_upd, n_st);
            if (COUNT) ++n_ld;
            vp = ((z + 1) & 7) ? vp + 16 : vol4 + quad_index(a, pl, z + 1); // next plane of the brick, or the next brick layer
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
                                            ctx->tab4, ctx->zexit);
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

// planned sweep: plan -> states of the general items -> general items || stream items
static int launch_integrate_planned(kfb_ctx *ctx, IntegrateArgs &a, int planes, uint64_t *n_updated)
{
    a.zchunk = KFB_PLAN_ZCHUNK;
    if (const char *e = getenv("KFB_PLAN_ZCHUNK")) { const int v = atoi(e); if (v >= 2 && v <= 64) a.zchunk = v; }
    a.nchunks = (a.ze - 1) / a.zchunk - a.zb / a.zchunk + 1;
    a.npx = (a.X + KFB_PATCH_X - 1) / KFB_PATCH_X;
    a.npy = (a.Y + KFB_PATCH_Y - 1) / KFB_PATCH_Y;
    a.mask_words = (a.nchunks + 31) / 32;
    if (a.ze > 65535) { ctx->err = "volumes deeper than 65535 planes are not supported"; return KFB_ERR_UNSUPPORTED; }
    const size_t npatch = (size_t)a.npx * a.npy, ncell = npatch * a.nchunks;
    const size_t gcap = std::max<size_t>(4096, ncell / 4);
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_counts = 0, o_mask = 256, o_slot = o_mask + up(npatch * a.mask_words * 4), o_is = o_slot + up(ncell * 4),
                 o_ig = o_is + up(ncell * 8), o_rd = o_ig + up(ncell * 8), o_st = o_rd + up(gcap * 4), need = o_st + gcap * 192 * 8;
    if (need > ctx->plan_bytes)
    {
        KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->plan_buf) cudaFree(ctx->plan_buf);
        ctx->plan_buf = nullptr; ctx->plan_bytes = 0;
        KFB_CUDA(ctx, cudaMalloc(&ctx->plan_buf, need));
        ctx->plan_bytes = need;
        ctx->plan_clean_bytes = 0;
        KFB_CUDA(ctx, cudaMemsetAsync(ctx->plan_buf, 0, need, ctx->stream)); // the ready tags start from a known value
    }
    char *pb = (char *)ctx->plan_buf;
    a.plan_counts = (unsigned int *)(pb + o_counts);
    a.patch_mask = (unsigned int *)(pb + o_mask);
    a.slot_of = (unsigned int *)(pb + o_slot);
    a.items_stream = (uint2 *)(pb + o_is);
    a.items_general = (uint2 *)(pb + o_ig);
    a.gstates = (unsigned long long *)(pb + o_st);
    a.ready = (unsigned int *)(pb + o_rd);
    a.ready_tag = (unsigned int)(++ctx->integrate_seq);
    if (a.ready_tag == 0) a.ready_tag = (unsigned int)(++ctx->integrate_seq);
    a.err = ctx->dev_err_dev;
    a.gstate_cap = (int)std::min<size_t>(gcap, 0x7fffffff);
    // counters + masks start from zero: normally the previous call has cleared them on the second stream, behind
    // everything that read them (see the end of this function), and the frame's critical path only waits for that event
    if (ctx->plan_clean_bytes == o_slot) KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_plan_clean, 0));
    else KFB_CUDA(ctx, cudaMemsetAsync(pb, 0, o_slot, ctx->stream));
    ctx->plan_clean_bytes = 0;
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
    int gwarps = 4; // warps (items) per block of the general kernel (KFB_GEN_WARPS: tuning)
    if (const char *e = getenv("KFB_GEN_WARPS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4) gwarps = v; }
    if (!getenv("KFB_INTEGRATE_PERSISTENT"))
    {
        // about one warp per item, sized from the previous frame's counts (both kernels stride, so any grid is
        // correct): the block scheduler balances the two kernels over whatever the SMs have free
        const size_t hs = ctx->plan_hint_host ? (size_t)ctx->plan_hint_host[0] : 0, hg = ctx->plan_hint_host ? (size_t)ctx->plan_hint_host[1] : 0;
        const size_t cap = std::max<size_t>((ncell + 3) / 4, 1);
        gs = (int)std::min<size_t>(std::max<size_t>((hs + hs / 4 + 3) / 4, (size_t)gs), cap);
        gg = (int)std::min<size_t>(std::max<size_t>((hg + hg / 4 + gwarps - 1) / gwarps, (size_t)gg), cap * (4 / gwarps));
    }
    const int gthreads = 32 * gwarps;
    cudaStream_t gstr = ctx->stream;
    if (overlap)
    {
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_ifork, ctx->stream));
        KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->istream, ctx->ev_ifork, 0));
        gstr = ctx->istream;
    }
    // the running sums of the general items: walked by the first blocks of the general kernel itself (fused, default)
    // or by a kernel of their own in front of it (KFB_INTEGRATE_SPLITSTATES=1)
    a.producer_blocks = 0;
    a.gen_prefetch = getenv("KFB_GEN_NOPREFETCH") ? 0 : 1;
    if (getenv("KFB_INTEGRATE_SPLITSTATES"))
    {
        integrate_states_kernel<<<(unsigned)((npatch + 3) / 4), 128, 0, gstr>>>(a);
        KFB_LAUNCH_CHECK(ctx);
    }
    else
    {
        a.producer_blocks = (int)((npatch + gwarps - 1) / gwarps);
        gg += a.producer_blocks;
    }
    if (n_updated)
    {
        if (smem) integrate_general_kernel<true, true, KFB_GEN_MINB><<<gg, gthreads, 0, gstr>>>(a);
        else integrate_general_kernel<true, false, KFB_GEN_MINB><<<gg, gthreads, 0, gstr>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        if (smem) integrate_stream_kernel<true, true><<<gs, 128, 0, ctx->stream>>>(a);
        else integrate_stream_kernel<true, false><<<gs, 128, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
    }
    else
    {
        const int minb = getenv("KFB_GEN_MINB") ? atoi(getenv("KFB_GEN_MINB")) : KFB_GEN_MINB; // register budget variant (tuning)
        if (!smem) integrate_general_kernel<false, false, KFB_GEN_MINB><<<gg, gthreads, 0, gstr>>>(a);
        else if (minb == 5) integrate_general_kernel<false, true, 5><<<gg, gthreads, 0, gstr>>>(a);
        else if (minb == 6) integrate_general_kernel<false, true, 6><<<gg, gthreads, 0, gstr>>>(a);
        else integrate_general_kernel<false, true, 8><<<gg, gthreads, 0, gstr>>>(a);
        KFB_LAUNCH_CHECK(ctx);
        if (smem) integrate_stream_kernel<false, true><<<gs, 128, 0, ctx->stream>>>(a);
        else integrate_stream_kernel<false, false><<<gs, 128, 0, ctx->stream>>>(a);
        KFB_LAUNCH_CHECK(ctx);
    }
    if (overlap)
    {
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_sweep_main, ctx->stream)); // the stream items are done (read by the clearing below)
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_ijoin, ctx->istream));
        KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ijoin, 0));
    }
    if (ctx->profiling) cudaEventRecord(ctx->events[61], ctx->stream);
    // item counts of this frame -> pinned host memory, read (one frame late, unsynchronised) as the next grid hint;
    // on the second stream behind the join event, so that nothing on the frame's critical path waits for the copy
    if (ctx->plan_hint_host)
        KFB_CUDA(ctx, cudaMemcpyAsync(ctx->plan_hint_host, a.plan_counts, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, overlap ? ctx->istream : ctx->stream));
    // the last raycast's tile order (kfb_raycast.cu: flush_ray_order): behind this call's plan kernel, hence behind the ICP
    if (const int rco = flush_ray_order(ctx, overlap ? ctx->ev_ifork : nullptr)) return rco;
    if (overlap && !n_updated)
    {
        // both sweep kernels and the hint copy are behind us on the second stream: clear the plan's counters and masks
        // for the next call there, off the critical path
        KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->istream, ctx->ev_sweep_main, 0));
        KFB_CUDA(ctx, cudaMemsetAsync(pb, 0, o_slot, ctx->istream));
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_plan_clean, ctx->istream));
        ctx->plan_clean_bytes = o_slot;
    }
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
    a.zb = ctx->z0 < 1 ? 1 : ctx->z0;
    a.ze = ctx->z1;
    a.pose = make_pose(vol2cam12);
    a.vsx = ctx->voxel_size[0]; a.vsy = ctx->voxel_size[1]; a.vsz = ctx->voxel_size[2];
    a.trunc = ctx->p.volu_trun_dist;
    a.fx = k.fx; a.fy = k.fy; a.cx = k.cx; a.cy = k.cy;
    a.w = k.w; a.h = k.h;
    a.wtab = ctx->wtab;
    a.zexit = ctx->zexit;
    a.max_weight = ctx->p.tsdf_max_weight;
    a.no_fastpath = getenv("KFB_INTEGRATE_NOFAST") ? 1 : 0;
    a.use_jump = getenv("KFB_INTEGRATE_NOJUMP") ? 0 : 1;
    // measured on B200 (slabs of 1024^3 and 2048^3, profiles/r02_experiments.md): the jump in front of a slab pays from
    // a few hundred planes on (thresholds 160 and 320 equal, 640 slower by 4-8 us per sweep at 1024^3); warps whose
    // four x or y values sit on both sides of zero replay in plain runs inside jump_fma
    a.jump_min = getenv("KFB_INTEGRATE_JUMPMIN") ? atoi(getenv("KFB_INTEGRATE_JUMPMIN")) : 256;
    a.bricks = ctx->bricks;
    a.bdirty = ctx->bdirty;
    a.dirty_tag = ctx->bdirty_tag;
    a.bx = ctx->bdim[0]; a.by = ctx->bdim[1]; a.bz = ctx->bdim[2]; a.bz0 = ctx->bz0;
    a.counter = ctx->counters;
    make_cull_planes(ctx, a, a.cull);
    a.Sx = a.vsx * a.pose.R.m[2]; a.Sy = a.vsx * a.pose.R.m[5]; a.Sz = a.vsx * a.pose.R.m[8];
    a.invSz = a.Sz > 1e-6f ? 1.f / a.Sz : 0.f;
    {
        const double M = fabs(a.pose.t[0]) + fabs(a.pose.t[1]) + fabs(a.pose.t[2]) + (double)ctx->p.volu_range[0] + ctx->p.volu_range[1] +
                         ctx->p.volu_range[2] + (double)a.vsx * ctx->p.volu_dims[2] * 1.01;
        a.driftE = (float)(((double)ctx->p.volu_dims[2] + 16.0) * 1.2e-7 * M + 4e-6 * M); // as in make_cull_planes + float model evaluation
    }
    if (getenv("KFB_INTEGRATE_NOCULL") || getenv("KFB_INTEGRATE_NOOCC")) a.Sz = 0.f; // no occlusion cut, no stream items
    if (getenv("KFB_INTEGRATE_NOCULL"))
        for (int c = 0; c < KFB_NCULL; ++c) a.cull[c].kind = 3;
    a.tab4 = ctx->tab4;
    a.zsparse = ctx->zsparse;
}

int launch_integrate(kfb_ctx *ctx, const float vol2cam12[12], uint64_t *n_updated)
{
    if (ctx->profiling) cudaEventRecord(ctx->events[56], ctx->stream); // whole call: 56 .. 57
    IntegrateArgs a;
    ctx->bdirty_tag = ctx->bdirty_tag == 0x7fffffff ? 1 : ctx->bdirty_tag + 1;
    fill_integrate_args(ctx, vol2cam12, a);
    const int planes = a.ze - a.zb;
    if (planes <= 0) return KFB_OK;
    const int rcp = launch_integrate_planned(ctx, a, planes, n_updated);
    if (rcp) return rcp;
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_tables_free, ctx->stream)); // the next frame's tables may now be built
    const int rcd = launch_brick_distance(ctx, false);
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
// full rescan (after kfb_upload_volume): same marking rule as the integrate kernel; a thread looks at one 16-byte
// quad of the brick-major volume
__global__ void rebuild_bricks_kernel(const IntegrateArgs a, size_t nquads, int Z)
{
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < nquads; q += (size_t)gridDim.x * blockDim.x)
    {
        const uint4 o = __ldg(reinterpret_cast<const uint4 *>(a.vol) + q);
        if (!((o.x | o.y | o.z | o.w) & 0x8000u)) continue;
        const size_t brick = q >> 7;
        const int in = (int)(q & 127), bxy = a.bx * a.by;
        const int bzi = (int)(brick / bxy), r = (int)(brick - (size_t)bzi * bxy);
        const int x0 = ((r % a.bx) << 3) + ((in & 1) << 2), y = ((r / a.bx) << 3) + ((in >> 1) & 7), z = ((bzi + a.bz0) << 3) + (in >> 4);
        if (x0 < a.X && y < a.Y && z < Z) mark_bricks(a.bricks, a.bdirty, a.dirty_tag, a.bx, a.by, a.bz, a.bz0, x0, y, z);
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
    ctx->bdirty_tag = ctx->bdirty_tag == 0x7fffffff ? 1 : ctx->bdirty_tag + 1;
    a.dirty_tag = ctx->bdirty_tag;
    a.bx = ctx->bdim[0]; a.by = ctx->bdim[1]; a.bz = ctx->bdim[2]; a.bz0 = ctx->bz0;
    rebuild_bricks_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(a, ctx->vol_voxels / 4, ctx->p.volu_dims[2]);
    KFB_LAUNCH_CHECK(ctx);
    return launch_brick_distance(ctx, true); // whatever the flags did: the map is rebuilt
}

// ---- brick distance map ----------------------------------------------------------------------------
// bdist[b] = min(KFB_BDIST_CAP, Chebyshev distance in bricks from b to the nearest active brick), 0 for an
// active brick.  Three separable passes: d(b) = min over offsets j along the axis of max(src(b + j), |j|).
// Every pass returns at once unless a brick flag has flipped in the sweep it follows (*dirty == that call's tag);
// flags only ever flip 0 -> 1 between resets, so a clean map stays exact.
__global__ void brick_distance_kernel(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int bx, int by, int bz, int axis,
                                      const int *__restrict__ dirty, int dirty_tag, int from_flags)
{
    if (*dirty != dirty_tag) return; // no flag flipped in the sweep this map follows
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = bx * by * bz;
    if (i >= n) return;
    const int x = i % bx, y = (i / bx) % by, z = i / (bx * by);
    const int pos = axis == 0 ? x : (axis == 1 ? y : z);
    const int len = axis == 0 ? bx : (axis == 1 ? by : bz);
    const int stride = axis == 0 ? 1 : (axis == 1 ? bx : bx * by);
    // taps in rings |j| = 0, 1, 2, ...: a tap at distance r contributes max(v, r) >= r, so the search ends as soon as
    // r reaches the best value found (near surfaces after a few taps; only far bricks look at all 31)
    int best = src[i];
    if (from_flags) best = best ? 0 : KFB_BDIST_CAP;
    best = min(best, KFB_BDIST_CAP);
    for (int r = 1; r < best; ++r)
    {
        int lo = pos - r >= 0 ? src[i - r * stride] : KFB_BDIST_CAP;
        int hi = pos + r < len ? src[i + r * stride] : KFB_BDIST_CAP;
        if (from_flags) { lo = (pos - r >= 0 && lo) ? 0 : KFB_BDIST_CAP; hi = (pos + r < len && hi) ? 0 : KFB_BDIST_CAP; }
        best = min(best, max(min(lo, hi), r));
    }
    dst[i] = (uint8_t)best;
}

int launch_brick_distance(kfb_ctx *ctx, bool force)
{
    // the passes run when the sweep of this call left its tag in *bdirty; `force`: compare with whatever is there
    const int *dirty = ctx->bdirty;
    const int tag = ctx->bdirty_tag;
    if (force) KFB_CUDA(ctx, cudaMemcpyAsync(ctx->bdirty, &ctx->bdirty_tag, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int n = ctx->bdim[0] * ctx->bdim[1] * ctx->bdim[2];
    const int blocks = (n + 255) / 256;
    // (measured and dropped in round 2: the x and y passes of a brick layer fused in shared memory, one block per layer
    // -- 15.4 us against 2 x 7.3 us: too few blocks; and all three passes in one cooperative launch: +25 us per call)
    brick_distance_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bricks, ctx->bdist_tmp, ctx->bdim[0], ctx->bdim[1], ctx->bdim[2], 0, dirty, tag, 1);
    KFB_LAUNCH_CHECK(ctx);
    brick_distance_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bdist_tmp, ctx->bdist_tmp2, ctx->bdim[0], ctx->bdim[1], ctx->bdim[2], 1, dirty, tag, 0);
    KFB_LAUNCH_CHECK(ctx);
    brick_distance_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->bdist_tmp2, ctx->bdist, ctx->bdim[0], ctx->bdim[1], ctx->bdim[2], 2, dirty, tag, 0);
    KFB_LAUNCH_CHECK(ctx);
    return KFB_OK;
}

// ---- work histogram over planes (slab balancing for sharded volumes) -----------------------------------------
// Work per plane from the sweep's own plan: the plan kernel runs for the WHOLE volume (it needs the depth tables and
// the pose, not the voxels), and every item adds its 32 voxel quads per plane to the planes it covers -- once for a
// stream item, KFB_HIST_GENERAL_WEIGHT times for a general item: about 4 for the sweep itself (measured cost ratio
// per quad at 512^3) plus the slab's share of the raycast, whose fetches and normals happen on the same planes
// (KFB_HIST_GENERAL_WEIGHT=<n> in the environment overrides; 8 GPUs / 1024^3 with weight 4: rank 0 at 0.08 + 0.03 ms,
// rank 7 at 0.14 + 0.21 ms).  (A count
// of the frustum alone left the busiest of two 2048^3 slabs 30 % behind the other, profiles/README.md.)
#define KFB_HIST_GENERAL_WEIGHT 10
#define KFB_HIST_RAY_UNITS 190 // raycast cost of one pixel in the unit of the histogram (a streamed quad): measured 0.144 ms per
                               // 307 200 rays against 0.103 ms per 6.8 M streamed + 10 x 3.5 M per-voxel quads at 512^3
#define KFB_HIST_RAY_PLANES 16 // planes in front of a pixel's surface over which that cost is spread (the near-surface march)
// raycast share of the work histogram: every pixel with a valid depth puts its cost on the planes just in front of the
// plane its surface point falls on (volume coordinates through cam2vol = the inverse of the rigid vol2cam)
__global__ void ray_histogram_kernel(const float *__restrict__ depth, int w, int h, float fx, float fy, float cx, float cy, Pose v2c, float vsz, int Z,
                                     int units, int *__restrict__ diff)
{
    const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
    if (u >= w || v >= h) return;
    const float d = depth[v * w + u];
    if (!(d > 0.f)) return;
    const float px = d * ((float)u - cx) / fx - v2c.t[0], py = d * ((float)v - cy) / fy - v2c.t[1], pz = d - v2c.t[2];
    const float zv = v2c.R.m[2] * px + v2c.R.m[5] * py + v2c.R.m[8] * pz; // third row of R^T
    const int zh = (int)floorf(zv / vsz);
    const int z1 = min(max(zh, 1), Z - 1), z0 = max(z1 - (KFB_HIST_RAY_PLANES - 1), 1);
    const int per = units / (z1 - z0 + 1);
    atomicAdd(diff + z0, per);
    atomicAdd(diff + z1 + 1, -per);
}

int launch_plane_histogram(kfb_ctx *ctx, const float vol2cam12[12], uint32_t *host_hist)
{
    const int Z = ctx->p.volu_dims[2];
    IntegrateArgs a;
    fill_integrate_args(ctx, vol2cam12, a);
    a.vol = nullptr;
    a.zb = 1; a.ze = Z;
    if (Z > 65535) { ctx->err = "volumes deeper than 65535 planes are not supported"; return KFB_ERR_UNSUPPORTED; }
    a.zchunk = 16;
    a.nchunks = (a.ze - 1) / a.zchunk - a.zb / a.zchunk + 1;
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
    // raycast share
    std::vector<int> rdiff((size_t)Z + 2, 0);
    if (e == cudaSuccess && !getenv("KFB_HIST_NORAY"))
    {
        int *dd = reinterpret_cast<int *>(pb + o_mask); // the plan has run: its patch masks are not needed any more
        if ((size_t)(Z + 2) * sizeof(int) <= o_slot - o_mask)
        {
            const Intr &k = ctx->L[0].k;
            e = cudaMemsetAsync(dd, 0, (size_t)(Z + 2) * sizeof(int), ctx->stream);
            if (e == cudaSuccess)
            {
                dim3 rb(32, 8), rg((k.w + 31) / 32, (k.h + 7) / 8);
                ray_histogram_kernel<<<rg, rb, 0, ctx->stream>>>(ctx->L[0].depth, k.w, k.h, k.fx, k.fy, k.cx, k.cy, a.pose, a.vsz, Z, KFB_HIST_RAY_UNITS, dd);
                ctx->launches++;
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaMemcpyAsync(rdiff.data(), dd, (size_t)(Z + 2) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        }
    }
    std::vector<uint2> items[2];
    for (int t = 0; t < 2 && e == cudaSuccess; ++t)
    {
        items[t].resize(counts[t]);
        if (counts[t]) e = cudaMemcpy(items[t].data(), t ? a.items_general : a.items_stream, (size_t)counts[t] * sizeof(uint2), cudaMemcpyDeviceToHost);
    }
    cudaFree(pb);
    KFB_CUDA(ctx, e);
    int gweight = KFB_HIST_GENERAL_WEIGHT;
    if (const char *e = getenv("KFB_HIST_GENERAL_WEIGHT")) { const int v = atoi(e); if (v >= 1 && v <= 1000) gweight = v; }
    std::vector<long long> diff((size_t)Z + 2, 0);
    for (int t = 0; t < 2; ++t)
    {
        const long long wgt = 32 * (t ? gweight : 1);
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
        run += diff[z] + (long long)rdiff[z];
        host_hist[z] = (uint32_t)std::min<long long>(std::max<long long>(run, 0), 0xffffffffll);
    }
    return KFB_OK;
}

// ---- host <-> device volume copies in the reference's order (GpuMat download / upload of the tests and checkpoints) ----
// lin: X * Y * (z1 - z0) packed voxels, x fastest; a thread moves four consecutive x voxels
__global__ void volume_relayout_kernel(uint32_t *__restrict__ vol, uint32_t *__restrict__ lin, int X, int Y, int z0, int z1, int bx, int by, int bz0,
                                       int to_bricks)
{
    const size_t nq = ((size_t)X * Y * (size_t)(z1 - z0)) >> 2;
    for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < nq; q += (size_t)gridDim.x * blockDim.x)
    {
        const size_t i = q << 2;
        const int x = (int)(i % X), y = (int)((i / X) % Y), z = z0 + (int)(i / ((size_t)X * Y));
        uint4 *b = reinterpret_cast<uint4 *>(vol + vol_index(bx, by, bz0, x, y, z));
        uint4 *l = reinterpret_cast<uint4 *>(lin + i);
        if (to_bricks) *b = *l;
        else *l = *b;
    }
}

int launch_volume_copy(kfb_ctx *ctx, int16_t *host_pairs, int to_device)
{
    uint32_t *lin = nullptr;
    const size_t bytes = ctx->vol_logical * sizeof(uint32_t);
    KFB_CUDA(ctx, cudaMalloc(&lin, bytes));
    cudaError_t e = cudaSuccess;
    if (to_device) e = cudaMemcpyAsync(lin, host_pairs, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
    {
        volume_relayout_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(ctx->vol, lin, ctx->p.volu_dims[0], ctx->p.volu_dims[1], ctx->z0, ctx->z1,
                                                                           ctx->bdim[0], ctx->bdim[1], ctx->bz0, to_device);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && !to_device) e = cudaMemcpyAsync(host_pairs, lin, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(lin);
    KFB_CUDA(ctx, e);
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
