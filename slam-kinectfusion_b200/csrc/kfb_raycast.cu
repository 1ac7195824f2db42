// kfb_raycast.cu -- TSDF raycast (replaces kf::device::raycast, kfusion/src/tsdf_volume.cu:120-273)
// and the render kernels (kf::device::renderPhong / renderNormals, image_process.cu:137-221).
//
// Semantics are the reference's, rounding for rounding: one-voxel steps,
// nearest-neighbour sample for the sign test, first +->- pair => vertex at
// ray_len -/+ vs*cur/(cur-next) (the reference's sign quirk behind compat_raycast_ts_sign),
// normal = normalised central difference of six trilinear fetches, outputs rotated back to the
// camera frame, explicit zeros on a miss.  The running sums `nextp += dir*voxel_size` and
// `ray_len += step` are replayed exactly, step for step.  Structure is new:
//   * a warp owns an 8x4 pixel tile (coherent rays) and marches it cooperatively: when no ray needs a fetch the
//     warp skips the minimum of the rays' safe step counts, otherwise four steps are classified and their loads
//     issued before the first sign test; candidate hits are parked and their normals computed after the march;
//     the two coarser levels of the model pyramid are written from the tile by register shuffles;
//   * empty-space skipping that cannot change a result: a terminal event (hit or back-face stop) needs
//     one negative and one positive sample in consecutive steps, and consecutive samples are at most two
//     voxels apart per axis, so a sample whose 8^3 brick has no negative voxel within two voxels (the byte
//     map kfb_integrate.cu maintains) can take part in no event.  Such samples are not fetched (they count
//     as NaN, which the reference's `isnan(next)` test already skips); a Chebyshev distance map over the
//     bricks says how many further steps provably stay clear of every such brick, and those steps are run
//     with the running-sum instructions only (two packed FFMA2 per step);
//   * tiles are marched in the order of their cost in the previous frame (the longest rays first);
//   * index rounding by magic-number add instead of F2I (the quarter-rate conversion pipe);
//   * z-slab mode (sharded volumes): samples outside the stored planes are "not mine"
//     (NaN), events are evaluated only when the `next` sample's voxel plane is owned by this slab, and the
//     ray length of the first terminal event is written as an order-preserving key for the cross-slab
//     composite (kfb_composite_mask).
#include "kfb_common.cuh"

namespace kfb
{

struct RaycastArgs
{
    const uint32_t *vol;
    int X, Y, Z;          // global dims
    int zs0, zs1;         // stored planes [zs0, zs1)
    int zo0, zo1;         // owned planes  [zo0, zo1)
    Pose pose;            // cam2vol
    Mat3 rinv;
    Intr k;
    float vs[3], vsinv[3], gd[3], range[3];
    float step_len;
    float4 *vmap, *nmap;
    float *key;
    int ts_sign_compat;
    const uint8_t *bdist;
    int bx, by, bz, bz0;
    int sparse_maps;      // slab pushing into a peer's staging: maps are written where the ray had an event only
    int fuse_pyramid;     // write levels 1 and 2 of the model maps from the warp tile (single-GPU, aligned image sizes)
    float4 *pyr_v[2], *pyr_n[2];
    // tile scheduling: block b marches tile order[b] (most expensive tiles of the previous frame first) and leaves
    // its own cost (SM cycles) for the next frame's order
    const unsigned int *order;
    unsigned int *cost;
    int tiles_x;
};

#define KFB_RC_MAGIC_F 12582912.0f
#define KFB_RC_MAGIC_I 0x4B400000
#define KFB_QNAN __int_as_float(0x7fffffff)

// __float2int_rn for |v| < 2^22 (larger magnitudes and NaN come out far outside any volume)
__device__ __forceinline__ int rn_magic(float v) { return __float_as_int(__fadd_rn(v, KFB_RC_MAGIC_F)) - KFB_RC_MAGIC_I; }
// (float)i for |i| < 2^22
__device__ __forceinline__ float i2f_magic(int i) { return __fsub_rn(__int_as_float(KFB_RC_MAGIC_I + i), KFB_RC_MAGIC_F); }

// MODE RC_WHOLE: the context stores the whole volume (zs0 = zo0 = bz0 = 0, zs1 = zo1 = Z); the slab tests fold away
// MODE RC_SLAB:  z-slab (stored planes [zs0, zs1), owned planes [zo0, zo1))
// Voxels are read from the brick-major volume (kfb_common.cuh: vol_index; the brick grid is the distance map's).
enum { RC_WHOLE = 0, RC_SLAB = 1 };
template <int MODE>
__device__ __forceinline__ size_t vox_index(const RaycastArgs &a, int x, int y, int z)
{
    return vol_index(a.bx, a.by, MODE == RC_SLAB ? a.bz0 : 0, x, y, z);
}
template <int MODE>
__device__ __forceinline__ float vox_tsdf(const RaycastArgs &a, int x, int y, int z)
{
    if (MODE == RC_SLAB) z = min(max(z, a.zs0), a.zs1 - 1); // slab mode: never read outside the stored planes
    const short s = __ldg(reinterpret_cast<const short *>(a.vol + vox_index<MODE>(a, x, y, z))); // low half = tsdf
    return __fmul_rn((float)s, KFB_DIVSHORTMAX);
}
// interpolate (tsdf_volume.cu:137-161)
template <int MODE>
__device__ float interp(const RaycastArgs &a, float fx, float fy, float fz)
{
    const int gx = __float2int_rd(fx), gy = __float2int_rd(fy), gz = __float2int_rd(fz);
    if (gx < 0 || gx >= a.X - 1 || gy < 0 || gy >= a.Y - 1 || gz < 0 || gz >= a.Z - 1) return KFB_QNAN;
    const float fa = __fsub_rn(fx, (float)gx), fb = __fsub_rn(fy, (float)gy), fc = __fsub_rn(fz, (float)gz);
    const float a1 = __fsub_rn(1.f, fa), b1 = __fsub_rn(1.f, fb), c1 = __fsub_rn(1.f, fc);
    const float v000 = vox_tsdf<MODE>(a, gx, gy, gz), v001 = vox_tsdf<MODE>(a, gx, gy, gz + 1);
    const float v010 = vox_tsdf<MODE>(a, gx, gy + 1, gz), v011 = vox_tsdf<MODE>(a, gx, gy + 1, gz + 1);
    const float v100 = vox_tsdf<MODE>(a, gx + 1, gy, gz), v101 = vox_tsdf<MODE>(a, gx + 1, gy, gz + 1);
    const float v110 = vox_tsdf<MODE>(a, gx + 1, gy + 1, gz), v111 = vox_tsdf<MODE>(a, gx + 1, gy + 1, gz + 1);
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

struct RaySkip // per-ray constants of the skip computation, voxel units
{
    float sgn[3];   // sign of the per-step move along each axis (+1 / -1)
    float inv[3];   // 1 / |move per step|  (1e30 when the ray does not move along the axis)
};

enum { ST_FETCH = 0, ST_NAN = 1, ST_LEAVE = 2 };
#define KFB_RC_BATCH 4

// raycasthelper::voxel2tsdf (tsdf_volume.cu:178-191) with brick / slab knowledge, split into "where" and
// "load".  ST_FETCH: the sample must be read (*addr).  ST_NAN: the sample is NaN for the march -- outside the
// volume interior, outside this slab's stored planes, or in a brick that cannot take part in an event --
// and `cnt` further steps are guaranteed to be NaN for the same reason.  ST_LEAVE: the ray moves away from
// the stored planes for good.  `own` = the sample's voxel plane belongs to this slab.
template <int MODE>
__device__ __forceinline__ int classify(const RaycastArgs &a, const RaySkip &rs, float px, float py, float pz, bool &own,
                                        int &cnt, const short *&addr)
{
    constexpr bool SLAB = MODE == RC_SLAB;
    const float qx = __fmul_rn(px, a.vsinv[0]), qy = __fmul_rn(py, a.vsinv[1]), qz = __fmul_rn(pz, a.vsinv[2]);
    const int x = rn_magic(qx), y = rn_magic(qy), z = rn_magic(qz);
    own = false; cnt = 0; addr = nullptr;
    if ((unsigned)(x - 1) >= (unsigned)(a.X - 2) || (unsigned)(y - 1) >= (unsigned)(a.Y - 2) || (unsigned)(z - 1) >= (unsigned)(a.Z - 2))
        return ST_NAN;
    if (SLAB && (z < a.zs0 || z >= a.zs1))
    {
        // outside the stored planes: steps until the sample can reach them (1.5 voxels of margin for the
        // drift of the replayed running sum over a long run)
        const bool below = z < a.zs0;
        const bool toward = below ? rs.sgn[2] > 0.f : rs.sgn[2] < 0.f;
        if (!toward) return ST_LEAVE;
        const float dist = below ? __fsub_rn((float)a.zs0 - 2.0f, qz) : __fsub_rn(qz, (float)a.zs1 + 1.0f);
        const float s = __fmul_rn(dist, rs.inv[2]);
        cnt = s > 1.f ? (int)fminf(s, 1e6f) - 1 : 0;
        return ST_NAN;
    }
    own = !SLAB || (z >= a.zo0 && z < a.zo1);
    // Chebyshev distance D (in bricks) to the nearest brick that may hold a negative voxel; 0 = such a brick.
    // Index moves by at most n + 1 per axis over n steps, an active brick is at least (D-1)*8 + 1 voxels
    // away along some axis, so the next (D-1)*8 - 1 samples cannot lie in one.
    const int D = __ldg(a.bdist + (unsigned)((((z >> 3) - (SLAB ? a.bz0 : 0)) * a.by + (y >> 3)) * a.bx + (x >> 3))); // < 2^25 bricks
    if (D == 0)
    {
        addr = reinterpret_cast<const short *>(a.vol + vox_index<MODE>(a, x, y, z)); // low half = tsdf
        return ST_FETCH;
    }
    cnt = max((D - 1) * 8 - 1, 0);
    return ST_NAN;
}
__device__ __forceinline__ float load_tsdf(const short *addr)
{
    const int s = __ldg(addr);
    return __fmul_rn(i2f_magic(s), KFB_DIVSHORTMAX);
}

// Model pyramid (device::resizePointsNormals, image_process.cu:95-135) from a warp's 8x4 pixel tile: the tile holds
// complete 2x2 and 4x4 pixel groups, so levels 1 and 2 come from register shuffles in the reference's summation order
// ((d00 + d01) + d10) + d11, * 0.25, zero when any vertex x is NaN.  All 32 lanes must call it (lane = ty * 8 + tx).
__device__ __forceinline__ void pyramid_from_tile(float4 vout, float4 nout, int x, int y, int tx, int ty, bool inside, int w0, float4 *const pyr_v[2],
                                                  float4 *const pyr_n[2])
{
    const unsigned FULL = 0xffffffffu;

        float4 v = vout, n = nout;
        int lw = w0, lx = x, ly = y;
#pragma unroll
        for (int l = 1; l <= 2; ++l)
        {
            const int dxl = 1 << (l - 1), dyl = 8 << (l - 1); // lane distance of the +x / +y neighbour at this level
            float4 v01, v10, v11, n01, n10, n11;
            v01.x = __shfl_down_sync(FULL, v.x, dxl); v01.y = __shfl_down_sync(FULL, v.y, dxl); v01.z = __shfl_down_sync(FULL, v.z, dxl);
            v10.x = __shfl_down_sync(FULL, v.x, dyl); v10.y = __shfl_down_sync(FULL, v.y, dyl); v10.z = __shfl_down_sync(FULL, v.z, dyl);
            v11.x = __shfl_down_sync(FULL, v.x, dxl + dyl); v11.y = __shfl_down_sync(FULL, v.y, dxl + dyl); v11.z = __shfl_down_sync(FULL, v.z, dxl + dyl);
            n01.x = __shfl_down_sync(FULL, n.x, dxl); n01.y = __shfl_down_sync(FULL, n.y, dxl); n01.z = __shfl_down_sync(FULL, n.z, dxl);
            n10.x = __shfl_down_sync(FULL, n.x, dyl); n10.y = __shfl_down_sync(FULL, n.y, dyl); n10.z = __shfl_down_sync(FULL, n.z, dyl);
            n11.x = __shfl_down_sync(FULL, n.x, dxl + dyl); n11.y = __shfl_down_sync(FULL, n.y, dxl + dyl); n11.z = __shfl_down_sync(FULL, n.z, dxl + dyl);
            float4 vo = make_float4(0.f, 0.f, 0.f, 0.f), no = vo;
            if (!isnan(__fmul_rn(__fmul_rn(__fmul_rn(v.x, v01.x), v10.x), v11.x)))
            {
                vo.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(v.x, v01.x), v10.x), v11.x), 0.25f);
                vo.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(v.y, v01.y), v10.y), v11.y), 0.25f);
                vo.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(v.z, v01.z), v10.z), v11.z), 0.25f);
                no.x = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n.x, n01.x), n10.x), n11.x), 0.25f);
                no.y = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n.y, n01.y), n10.y), n11.y), 0.25f);
                no.z = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(n.z, n01.z), n10.z), n11.z), 0.25f);
            }
            v = vo; n = no;
            lw >>= 1; lx >>= 1; ly >>= 1;
            const int mask = (1 << l) - 1;
            if (inside && ((tx & mask) == 0) && ((ty & mask) == 0))
            {
                pyr_v[l - 1][ly * lw + lx] = v;
                pyr_n[l - 1][ly * lw + lx] = n;
            }
        }
}

// The march is warp-cooperative: all rays of an 8x4 tile step together.  When no ray of the warp needs a
// fetch for its next sample, the warp skips the minimum of the rays' guaranteed-NaN step counts with the
// running-sum instructions only; otherwise KFB_RC_BATCH steps are classified and their loads issued before
// the first sign test (sample positions do not depend on fetched values).  Candidate hits are parked and
// their normals are computed after the march, when the warp has reconverged.
#ifndef KFB_RC_WARPS
#define KFB_RC_WARPS 1 // warps (8x4 pixel tiles, stacked in y) per block; 2 warps with 16 / 18 / 21 blocks per SM measured 155 / 154 / 159 us against 152
#endif
#ifndef KFB_RC_MINB
#define KFB_RC_MINB 32 // launch-bounds hint: 32 one-warp blocks per SM (the hardware limit) => at most 64 registers
#endif
template <int MODE>
__global__ void __launch_bounds__(32 * KFB_RC_WARPS, KFB_RC_MINB) raycast_kernel(const RaycastArgs a)
{
    constexpr bool SLAB = MODE == RC_SLAB;
    const unsigned FULL = 0xffffffffu;
    // block = warp = 8x4 pixel tile (rays of very different length share nothing: fine-grained scheduling)
    const long long t_begin = clock64();
    const int tile = a.order ? (int)__ldg(a.order + blockIdx.x) : (int)blockIdx.x;
    const int x = (tile % a.tiles_x) * 8 + threadIdx.x;
    const int y = (tile / a.tiles_x) * 4 + threadIdx.y;
    const bool inside = x < a.k.w && y < a.k.h;
    const int pix = y * a.k.w + x;
    float4 vout = make_float4(0.f, 0.f, 0.f, 0.f), nout = vout;
    float key = __int_as_float(0x7f800000); // +inf = no event

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
    bool marching = inside && !(ray_len >= tfar);

    RaySkip rs;
    {
        // move per step in voxel units: dir * vs * (1/vs), slightly overestimated so step counts err low
        const float mx = fabsf(dx) * a.vs[0] * a.vsinv[0] * 1.0001f, my = fabsf(dy) * a.vs[1] * a.vsinv[1] * 1.0001f,
                    mz = fabsf(dz) * a.vs[2] * a.vsinv[2] * 1.0001f;
        rs.sgn[0] = dx < 0.f ? -1.f : 1.f; rs.sgn[1] = dy < 0.f ? -1.f : 1.f; rs.sgn[2] = dz < 0.f ? -1.f : 1.f;
        rs.inv[0] = mx > 1e-30f ? 1.f / mx : 1e30f;
        rs.inv[1] = my > 1e-30f ? 1.f / my : 1e30f;
        rs.inv[2] = mz > 1e-30f ? 1.f / mz : 1e30f;
    }
    const float rstep = 1.f / a.step_len;
    float nx = 0.f, ny = 0.f, nz = 0.f, tnext = KFB_QNAN;
    if (marching)
    {
        ray_len = __fadd_rn(ray_len, a.step_len);
        nx = __fmaf_rn(dx, ray_len, ox); ny = __fmaf_rn(dy, ray_len, oy); nz = __fmaf_rn(dz, ray_len, oz);
        bool own; int cnt; const short *addr;
        const int st = classify<MODE>(a, rs, nx, ny, nz, own, cnt, addr);
        if (st == ST_FETCH) tnext = load_tsdf(addr);
        if (st == ST_LEAVE) marching = false;
    }
    // parked candidate hit
    bool pend = false;
    float c_tcur = 0.f, c_tnext = 0.f, c_len = 0.f;

    for (;;)
    {
        // ---- march ----------------------------------------------------------------------------------
        for (;;)
        {
            const bool alive = marching && ray_len < tfar;
            if (!__any_sync(FULL, alive)) break;
            float qx[KFB_RC_BATCH], qy[KFB_RC_BATCH], qz[KFB_RC_BATCH];
            int st[KFB_RC_BATCH], cnt0 = 0;
            bool own[KFB_RC_BATCH];
            const short *addr[KFB_RC_BATCH];
            qx[0] = __fmaf_rn(dx, a.vs[0], nx); qy[0] = __fmaf_rn(dy, a.vs[1], ny); qz[0] = __fmaf_rn(dz, a.vs[2], nz);
            st[0] = ST_NAN; own[0] = false; addr[0] = nullptr;
            if (alive) st[0] = classify<MODE>(a, rs, qx[0], qy[0], qz[0], own[0], cnt0, addr[0]);
            if (__all_sync(FULL, !alive || st[0] != ST_FETCH))
            {
                // no ray of the warp needs this sample: it is NaN for all; then skip what every ray can skip.
                // A ray's count is also capped by its remaining steps to tfar (conservatively), so the skip
                // loop needs no per-step exit test; rays that are not marching keep their state untouched.
                int mine = 0x7fffffff;
                if (alive && st[0] == ST_NAN)
                {
                    const float rem = __fmul_rn(__fsub_rn(tfar, ray_len), rstep) - 3.f;
                    mine = min(cnt0, rem > 0.f ? (int)fminf(rem, 1e6f) : 0);
                }
                int nskip = __reduce_min_sync(FULL, mine);
                if (nskip == 0x7fffffff) nskip = 0;
                if (alive)
                {
                    if (st[0] == ST_LEAVE) marching = false;
                    nx = qx[0]; ny = qy[0]; nz = qz[0];
                    tnext = KFB_QNAN;
                    ray_len = __fadd_rn(ray_len, a.step_len);
                }
                if (marching && alive)
                {
                    {
                        // the four running sums as two packed FFMA2 per step (ray_len + step = fma(step, 1, ray_len), exactly).
                        // Also on the way to a far z-slab: the exact integer jump (kfb_common.cuh: jump_fma) was measured
                        // against this replay on slabs of 1024^3 and 2048^3 and lost for every run length a ray can have
                        // (1 800 steps: 264 us against 144 us per slab raycast; profiles/r02_experiments.md)
                        unsigned long long nxy = pack2(nx, ny), nzl = pack2(nz, ray_len);
                        const unsigned long long dxy = pack2(dx, dy), vxy = pack2(a.vs[0], a.vs[1]);
                        const unsigned long long dzl = pack2(dz, a.step_len), vz1 = pack2(a.vs[2], 1.0f);
#pragma unroll 4
                        for (int i = 0; i < nskip; ++i)
                        {
                            nxy = ffma2(dxy, vxy, nxy);
                            nzl = ffma2(dzl, vz1, nzl);
                        }
                        unpack2(nxy, nx, ny);
                        unpack2(nzl, nz, ray_len);
                    }
                }
                continue;
            }
            // ---- batch: classify the following steps too, issue all loads, then test in order ------------
#pragma unroll
            for (int b = 1; b < KFB_RC_BATCH; ++b)
            {
                qx[b] = __fmaf_rn(dx, a.vs[0], qx[b - 1]); qy[b] = __fmaf_rn(dy, a.vs[1], qy[b - 1]); qz[b] = __fmaf_rn(dz, a.vs[2], qz[b - 1]);
                st[b] = ST_NAN; own[b] = false; addr[b] = nullptr;
                int c;
                if (alive) st[b] = classify<MODE>(a, rs, qx[b], qy[b], qz[b], own[b], c, addr[b]);
            }
            float val[KFB_RC_BATCH];
#pragma unroll
            for (int b = 0; b < KFB_RC_BATCH; ++b) val[b] = (alive && st[b] == ST_FETCH) ? load_tsdf(addr[b]) : KFB_QNAN;
#pragma unroll
            for (int b = 0; b < KFB_RC_BATCH; ++b)
            {
                if (marching && ray_len < tfar)
                {
                    const float tcur = tnext;
                    tnext = val[b];
                    nx = qx[b]; ny = qy[b]; nz = qz[b];
                    if (st[b] == ST_LEAVE) marching = false;
                    else if (own[b] && !isnan(tnext))
                    {
                        if (tcur < 0.f && tnext > 0.f)
                        {
                            key = ray_len; // back face: stop, no hit (tsdf_volume.cu:242-243)
                            marching = false;
                        }
                        else if (tcur > 0.f && tnext < 0.f)
                        {
                            pend = true; // normal is computed after the march (may resume if it is NaN)
                            c_tcur = tcur; c_tnext = tnext; c_len = ray_len;
                            marching = false;
                        }
                    }
                    ray_len = __fadd_rn(ray_len, a.step_len);
                }
            }
        }
        // ---- parked candidates: vertex + normal (tsdf_volume.cu:244-259) -----------------------------------
        bool resumed = false;
        if (pend)
        {
            pend = false;
            const float q = rcp_fdividef(__fsub_rn(c_tcur, c_tnext));
            const float num = __fmul_rn(c_tcur, a.vs[0]);
            const float Ts = a.ts_sign_compat ? __fmaf_rn(q, -num, c_len) : __fmaf_rn(q, num, c_len);
            const float vx = __fmaf_rn(dx, Ts, ox), vy = __fmaf_rn(dy, Ts, oy), vz = __fmaf_rn(dz, Ts, oz);
            // compute_normal (tsdf_volume.cu:192-209)
            const float ux = __fmul_rn(vx, a.vsinv[0]), uy = __fmul_rn(vy, a.vsinv[1]), uz = __fmul_rn(vz, a.vsinv[2]);
            const float Fx1 = interp<MODE>(a, __fmul_rn(__fadd_rn(vx, a.gd[0]), a.vsinv[0]), uy, uz);
            const float Fx2 = interp<MODE>(a, __fmul_rn(__fsub_rn(vx, a.gd[0]), a.vsinv[0]), uy, uz);
            const float Fy1 = interp<MODE>(a, ux, __fmul_rn(__fadd_rn(vy, a.gd[1]), a.vsinv[1]), uz);
            const float Fy2 = interp<MODE>(a, ux, __fmul_rn(__fsub_rn(vy, a.gd[1]), a.vsinv[1]), uz);
            const float Fz1 = interp<MODE>(a, ux, uy, __fmul_rn(__fadd_rn(vz, a.gd[2]), a.vsinv[2]));
            const float Fz2 = interp<MODE>(a, ux, uy, __fmul_rn(__fsub_rn(vz, a.gd[2]), a.vsinv[2]));
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
                key = c_len;
            }
            else
            {
                marching = true; // the reference keeps marching after a NaN normal
                resumed = true;
            }
        }
        if (!__any_sync(FULL, resumed)) break;
    }
    if (inside)
    {
        // a slab rank that pushes into rank 0's staging sends the maps of the pixels that had an event only: the
        // composite reads the maps of a pixel's winning slab, and a slab without an event (+inf) never wins
        if (!(SLAB && a.sparse_maps) || key < __int_as_float(0x7f800000))
        {
            a.vmap[pix] = vout;
            a.nmap[pix] = nout;
        }
        a.key[pix] = key;
    }
    // ---- model pyramid as an epilogue of the warp tile (see pyramid_from_tile)
    if (a.fuse_pyramid) pyramid_from_tile(vout, nout, x, y, threadIdx.x, threadIdx.y & 3, inside, a.k.w, a.pyr_v, a.pyr_n);
    if (a.cost && threadIdx.x == 0 && threadIdx.y == 0) a.cost[tile] = (unsigned int)min((long long)0xffffffffll, clock64() - t_begin);
}

// Tile order for the next raycast: a counting sort of the tiles by the cost they just reported, most expensive
// first (one block, on a side stream).  Rays differ in length by two orders of magnitude (a ray that leaves the volume at once against
// one that creeps along a wall), the block scheduler hands tiles out in index order, and with image rows of floor
// at the bottom the expensive tiles used to come last: the kernel ended in a long tail of a few resident warps
// (achieved occupancy 32 % of a possible 50 %).  The camera moves slowly, so last frame's cost predicts this frame's.
#define KFB_RC_COST_SHIFT 11 // bucket = cycles / 2048, 256 buckets
__global__ void __launch_bounds__(1024) raycast_order_kernel(const unsigned int *__restrict__ cost, unsigned int *__restrict__ order, int ntiles)
{
    __shared__ unsigned int hist[256], base[256];
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x) atomicAdd(&hist[255 - min(255u, cost[i] >> KFB_RC_COST_SHIFT)], 1u);
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned int run = 0;
        for (int b = 0; b < 256; ++b) { base[b] = run; run += hist[b]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ntiles; i += blockDim.x)
    {
        const unsigned int b = 255 - min(255u, cost[i] >> KFB_RC_COST_SHIFT);
        order[atomicAdd(&base[b], 1u)] = (unsigned int)i;
    }
}

// cross-slab composite, step 2 (see include/kfb200.h): keep the payload only where this slab holds the
// winning (smallest) event key; ties cannot occur between slabs because a step has exactly one owner.
__global__ void composite_mask_kernel(const float *__restrict__ key, const float *__restrict__ min_key, float4 *__restrict__ vmap,
                                      float4 *__restrict__ nmap, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float k = key[i];
    if (!(k == min_key[i]) || k == __int_as_float(0x7f800000))
    {
        vmap[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        nmap[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

int launch_composite_mask(kfb_ctx *ctx, const float *min_key)
{
    const Intr &k = ctx->L[0].k;
    const int n = k.w * k.h;
    composite_mask_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->hit_t, min_key, ctx->L[0].v[ctx->prev], ctx->L[0].n[ctx->prev], n);
    KFB_LAUNCH_CHECK(ctx);
    return KFB_OK;
}

// ---- composite over NVLink peer memory (kfb_shard_composite) ---------------------------------------------------
struct ShardArgs
{
    const float *keys[16];
    const float4 *maps[16];                  // vertex map, then normal map (contiguous), of every rank's model slot
    const volatile unsigned long long *flag[16];
    float4 *out;                             // rank 0's own model maps
    int world, self, npix, w, h;
    int fuse_pyramid;                        // write levels 1 and 2 of the model maps from the tiles
    float4 *pyr_v[2], *pyr_n[2];
    unsigned long long seq;
    unsigned long long *err;                 // mapped host word: set to the frame's sequence number when a peer never signalled
};
__global__ void shard_signal_kernel(unsigned long long *flag, unsigned long long seq)
{
    // the raycast kernel before this launch has completed and pushed its results into rank 0's memory; the fence
    // orders them before the counter (rank 0's slot for this rank, a peer store as well)
    __threadfence_system();
    *(volatile unsigned long long *)flag = seq;
}
// One pixel per thread, a warp = an 8x4 pixel tile (as in the raycast): all peers' keys of a pixel are requested
// together (one NVLink round trip), then the winner's vertex and normal (a second one), and -- when the image sizes
// allow -- the two coarser levels of the model pyramid are written from the tile by register shuffles, exactly as the
// single-GPU raycast does in its epilogue (no resize launches afterwards).
__global__ void __launch_bounds__(256) shard_composite_kernel(const ShardArgs a)
{
    // wait until every slab of this frame has been raycast (flags live in the peers' memory, read over NVLink)
    if (threadIdx.x < a.world && threadIdx.x != a.self)
    {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while (*a.flag[threadIdx.x] < a.seq)
        {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t - t0 > 2000000000ull)
            {
                // 2 s: a peer died or stalled.  Leave rather than hang the GPU, and say so: the host finds the
                // word before it uses the composited maps (check_device_error) and fails the frame
                *(volatile unsigned long long *)a.err = a.seq;
                break;
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, tx = lane & 7, ty = lane >> 3;
    const int tiles_x = (a.w + 7) >> 3;
    const int tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int x = (tile % tiles_x) * 8 + tx, y = (tile / tiles_x) * 4 + ty;
    const bool inside = x < a.w && y < a.h;
    const int i = inside ? y * a.w + x : 0;
    float best = __int_as_float(0x7f800000);
    int win = -1;
    {
        float k[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) k[r] = (r < a.world && inside) ? __ldcv(a.keys[r] + i) : __int_as_float(0x7f800000);
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (k[r] < best) { best = k[r]; win = r; }
    }
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), n = v;
    if (win >= 0)
    {
        v = __ldcv(a.maps[win] + i);
        n = __ldcv(a.maps[win] + a.npix + i);
    }
    if (inside && win != a.self) { a.out[i] = v; a.out[a.npix + i] = n; }
    if (a.fuse_pyramid) pyramid_from_tile(v, n, x, y, tx, ty, inside, a.w, a.pyr_v, a.pyr_n);
}

int check_device_error(kfb_ctx *ctx)
{
    if (ctx->dev_err_host && *(volatile unsigned long long *)ctx->dev_err_host != 0ull)
    {
        const unsigned long long v = *(volatile unsigned long long *)ctx->dev_err_host;
        if (v >> 60)
            ctx->err = "integrate: the running sums of work item " + std::to_string(v & 0xffffffffull) + " never arrived; the volume is incomplete";
        else
            ctx->err = "shard composite: a peer's slab never arrived (frame " + std::to_string(v) + "); the model maps of that frame are incomplete";
        return KFB_ERR_TIMEOUT;
    }
    return KFB_OK;
}

int launch_shard_composite(kfb_ctx *ctx)
{
    const Intr &k = ctx->L[0].k;
    if (const int rce = check_device_error(ctx)) return rce;
    const unsigned long long seq = ++ctx->shard_seq;
    const size_t P = (size_t)k.w * k.h;
    if (ctx->shard_rank != 0)
    {
        shard_signal_kernel<<<1, 1, 0, ctx->stream>>>((unsigned long long *)ctx->peer_flag[0] + ctx->shard_rank, seq);
        KFB_LAUNCH_CHECK(ctx);
        return KFB_OK;
    }
    ShardArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < ctx->shard_world; ++r)
    {
        // rank 0's own slab sits in its raycast outputs, the peers' slabs in the slots they pushed into
        a.keys[r] = r == 0 ? ctx->hit_t : ctx->stage_keys + (size_t)r * P;
        a.maps[r] = r == 0 ? ctx->L[0].v[ctx->prev] : ctx->stage_maps + (size_t)r * 2 * P;
        a.flag[r] = (const volatile unsigned long long *)(ctx->shard_flag + r);
    }
    a.out = ctx->L[0].v[ctx->prev];
    a.world = ctx->shard_world; a.self = ctx->shard_rank; a.npix = k.w * k.h; a.seq = seq;
    a.w = k.w; a.h = k.h;
    a.err = ctx->dev_err_dev;
    a.fuse_pyramid = (ctx->levels == 3 && k.w % 8 == 0 && k.h % 4 == 0 && !getenv("KFB_RAYCAST_NOFUSE")) ? 1 : 0;
    for (int l = 1; l <= 2; ++l)
    {
        a.pyr_v[l - 1] = a.fuse_pyramid ? ctx->L[l].v[ctx->prev] : nullptr;
        a.pyr_n[l - 1] = a.fuse_pyramid ? ctx->L[l].n[ctx->prev] : nullptr;
    }
    const int tiles = ((k.w + 7) / 8) * ((k.h + 3) / 4);
    if (ctx->profiling) cudaEventRecord(ctx->events[52], ctx->stream); // composite (its wait for the slowest slab included): 52 .. 53
    shard_composite_kernel<<<(tiles + 7) / 8, 256, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    if (ctx->profiling) cudaEventRecord(ctx->events[53], ctx->stream);
    ctx->pyramid_fresh = a.fuse_pyramid;
    return KFB_OK;
}

int launch_raycast(kfb_ctx *ctx, const float cam2vol12[12], const float rinv9[9])
{
    RaycastArgs a;
    a.vol = ctx->vol;
    a.X = ctx->p.volu_dims[0]; a.Y = ctx->p.volu_dims[1]; a.Z = ctx->p.volu_dims[2];
    a.zs0 = ctx->z0; a.zs1 = ctx->z1;
    const bool slab = ctx->p.slab_z_end > ctx->p.slab_z_begin;
    a.zo0 = slab ? ctx->p.slab_z_begin : 0;
    a.zo1 = slab ? ctx->p.slab_z_end : a.Z;
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
    a.key = ctx->hit_t;
    a.sparse_maps = 0;
    if (ctx->shard_world > 0 && ctx->shard_rank != 0)
    {
        // attached slab rank: the kernel writes its keys and maps straight into its slot of rank 0's staging buffers
        // (posted NVLink stores from the epilogue; rank 0 then composites from its own memory)
        const size_t P = (size_t)a.k.w * a.k.h;
        a.key = (float *)ctx->peer_keys[0] + (size_t)ctx->shard_rank * P;
        a.vmap = (float4 *)ctx->peer_maps[0][0] + (size_t)ctx->shard_rank * 2 * P;
        a.nmap = a.vmap + P;
        a.sparse_maps = getenv("KFB_PUSH_DENSE") ? 0 : 1;
    }
    a.ts_sign_compat = ctx->p.compat_raycast_ts_sign;
    a.bdist = ctx->bdist;
    a.bx = ctx->bdim[0]; a.by = ctx->bdim[1]; a.bz = ctx->bdim[2]; a.bz0 = ctx->bz0;
    // the model pyramid rides along when the tile grid is exact and nothing has to be composited first
    a.fuse_pyramid = (!slab && ctx->levels == 3 && a.k.w % 8 == 0 && a.k.h % 4 == 0 && !getenv("KFB_RAYCAST_NOFUSE")) ? 1 : 0;
    for (int l = 1; l <= 2; ++l)
    {
        a.pyr_v[l - 1] = a.fuse_pyramid ? ctx->L[l].v[ctx->prev] : nullptr;
        a.pyr_n[l - 1] = a.fuse_pyramid ? ctx->L[l].n[ctx->prev] : nullptr;
    }
    ctx->pyramid_fresh = a.fuse_pyramid;
    a.tiles_x = (a.k.w + 7) / 8;
    const int ntiles = a.tiles_x * ((a.k.h + 3) / 4);
    const bool sorted = !getenv("KFB_RAYCAST_NOSORT");
    if (sorted && !ctx->ray_cost)
    {
        KFB_CUDA(ctx, cudaMalloc(&ctx->ray_cost, 2 * (size_t)ntiles * sizeof(unsigned int)));
        ctx->ray_order = ctx->ray_cost + ntiles;
        ctx->ray_order_valid = 0;
        KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ray_done, cudaEventDisableTiming));
        KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ray_order, cudaEventDisableTiming));
        KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_order_gate, cudaEventDisableTiming));
    }
    if (ctx->ray_order_pending) { if (const int rco = flush_ray_order(ctx, nullptr)) return rco; } // no integrate since the last raycast
    a.cost = sorted ? ctx->ray_cost : nullptr;
    a.order = (sorted && ctx->ray_order_valid) ? ctx->ray_order : nullptr;
    if (a.order) KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_ray_order, 0)); // long done: it ran behind the previous raycast
    dim3 block(8, 4), grid(ntiles);
    if (ctx->profiling) cudaEventRecord(ctx->events[58], ctx->stream);
    const bool whole = a.zs0 == 0 && a.zs1 == a.Z && a.zo0 == 0 && a.zo1 == a.Z && a.bz0 == 0 && !getenv("KFB_RAYCAST_SLABCODE");
    if (whole) raycast_kernel<RC_WHOLE><<<grid, block, 0, ctx->stream>>>(a);
    else raycast_kernel<RC_SLAB><<<grid, block, 0, ctx->stream>>>(a);
    KFB_LAUNCH_CHECK(ctx);
    if (ctx->profiling) cudaEventRecord(ctx->events[59], ctx->stream);
    if (sorted)
    {
        // The costs become next frame's order in a one-block kernel that nothing on the critical path waits for -- but
        // WHEN it runs matters: right behind this raycast it sat on one SM while the next frame's ICP kernel, which needs
        // a whole SM's registers on every SM for its co-resident grid, waited for it (12 us per frame).  It is launched
        // from the next kfb_integrate instead (flush_ray_order), behind that call's plan kernel -- i.e. behind the ICP --
        // on a stream of the lowest priority, by API calls the host makes while the sweep is already running.
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_ray_done, ctx->stream));
        ctx->ray_order_pending = 1;
        ctx->ray_order_tiles = ntiles;
    }
    return KFB_OK;
}

// the pending tile order of the last raycast, on the order stream: behind `gate` (an event already recorded on the main
// stream), or behind everything enqueued on the main stream so far
int flush_ray_order(kfb_ctx *ctx, cudaEvent_t gate)
{
    if (!ctx->ray_order_pending) return KFB_OK;
    if (!gate)
    {
        KFB_CUDA(ctx, cudaEventRecord(ctx->ev_order_gate, ctx->stream));
        gate = ctx->ev_order_gate;
    }
    KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->ostream, gate, 0));
    KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->ostream, ctx->ev_ray_done, 0));
    raycast_order_kernel<<<1, 1024, 0, ctx->ostream>>>(ctx->ray_cost, ctx->ray_order, ctx->ray_order_tiles);
    KFB_LAUNCH_CHECK(ctx);
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_ray_order, ctx->ostream));
    ctx->ray_order_valid = 1;
    ctx->ray_order_pending = 0;
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
