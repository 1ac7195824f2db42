/*
 * kf_oracle.c -- scalar/OpenMP CPU restatement of the reference hot path.
 * TEST INFRASTRUCTURE ONLY (see kf_oracle.h for the rules and parity status).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).  Build: see oracle/Makefile (gcc -O2 -fopenmp
 * -ffp-contract=off; the explicit fmaf() calls are the contraction nvcc applied
 * to the reference kernels for sm_100a).
 */
#include "kf_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#include <time.h>

/* Stand-in for MUFU.RCP behind __fdividef (device_utils.cuh:15-27 etc.). */
#define KFO_RCP(x) (1.0f / (x))
/* device_utils.cuh:5-7 */
#define KFO_DIVSHORTMAX 0.0000305185f
#define KFO_SHORTMAX 32767

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* cvt.rni.s32.f32 / cvt.rzi / cvt.rmi: saturating, NaN -> 0 */
static inline int f2i_sat(float r)
{
    if (isnan(r)) return 0;
    if (r >= 2147483648.0f) return INT32_MAX;
    if (r <= -2147483648.0f) return INT32_MIN;
    return (int)r;
}
static inline int f2i_rn(float v) { return f2i_sat(rintf(v)); }   /* __float2int_rn */
static inline int f2i_rd(float v) { return f2i_sat(floorf(v)); }  /* __float2int_rd */
static inline int f2i_rz(float v) { return f2i_sat(truncf(v)); }  /* static_cast<int> */

/* dot(v,v) as nvcc contracts x*x + y*y + z*z (device_types.hpp:238-241) */
static inline float dot3c(float ax, float ay, float az, float bx, float by, float bz)
{
    return fmaf(az, bz, fmaf(ax, bx, ay * by));
}
/* PoseR * float3 (device_types.hpp:138-143), R row-major m[9] */
static inline void rot3(const float *m, float x, float y, float z, float *o)
{
    o[0] = fmaf(z, m[2], fmaf(x, m[0], y * m[1]));
    o[1] = fmaf(z, m[5], fmaf(x, m[3], y * m[4]));
    o[2] = fmaf(z, m[8], fmaf(x, m[6], y * m[7]));
}
static inline void pose_R9(const float p[12], float R[9])
{
    R[0] = p[0]; R[1] = p[1]; R[2] = p[2];
    R[3] = p[4]; R[4] = p[5]; R[5] = p[6];
    R[6] = p[8]; R[7] = p[9]; R[8] = p[10];
}

/* ======================================================================
 * Front end
 * ====================================================================== */

/* types.hpp:18-28 */
void kfo_level_intrinsics(const kfo_intr *in, int level, kfo_intr *out)
{
    if (level == 0) { *out = *in; return; }
    const float s = powf(0.5f, (float)level);
    out->width = in->width >> level;
    out->height = in->height >> level;
    out->fx = in->fx * s;
    out->fy = in->fy * s;
    out->cx = (in->cx + 0.5f) * s - 0.5f;
    out->cy = (in->cy + 0.5f) * s - 0.5f;
}

static inline int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len)
    {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* cv::cuda::pyrDown as called at kinectfusion.cpp:55 (opencv_contrib cudawarping
 * pyr_down.cu, restated in SURVEY.md §10.1): vertical 5-tap first, then
 * horizontal, BORDER_REFLECT_101, each `sum + w*v` an FMA. */
void kfo_pyrdown(const float *src, int w, int h, float *dst, int dw, int dh)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y)
    {
        const int sy = 2 * y;
        const float *r0 = src + (size_t)reflect101(sy - 2, h) * w;
        const float *r1 = src + (size_t)reflect101(sy - 1, h) * w;
        const float *r2 = src + (size_t)reflect101(sy, h) * w;
        const float *r3 = src + (size_t)reflect101(sy + 1, h) * w;
        const float *r4 = src + (size_t)reflect101(sy + 2, h) * w;
        for (int x = 0; x < dw; ++x)
        {
            float col[5];
            for (int k = 0; k < 5; ++k)
            {
                const int c = reflect101(2 * x - 2 + k, w);
                float s = 0.0625f * r0[c];
                s = fmaf(0.25f, r1[c], s);
                s = fmaf(0.375f, r2[c], s);
                s = fmaf(0.25f, r3[c], s);
                s = fmaf(0.0625f, r4[c], s);
                col[k] = s;
            }
            float s = 0.0625f * col[0];
            s = fmaf(0.25f, col[1], s);
            s = fmaf(0.375f, col[2], s);
            s = fmaf(0.25f, col[3], s);
            s = fmaf(0.0625f, col[4], s);
            dst[(size_t)y * dw + x] = s;
        }
    }
}

/* cv::cuda::bilateralFilter as called at kinectfusion.cpp:60-64 (opencv_contrib
 * cudaimgproc bilateral_filter.cu).  Out-of-place (§9 Q1). */
void kfo_bilateral(const float *src, int w, int h, float *dst, int ksize, float sigma_color, float sigma_space)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = (ksize <= 0) ? (int)lrintf(sigma_space * 1.5f) : ksize / 2;
    if (radius < 1) radius = 1;
    const int r = radius; /* kernel_size = 2r+1; loop r = kernel_size/2 */
    const float r2 = (float)(r * r);
    const float ss = -0.5f / (sigma_space * sigma_space);
    const float sc = -0.5f / (sigma_color * sigma_color);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const float center = src[(size_t)y * w + x];
            float sum1 = 0.f, sum2 = 0.f;
            for (int cy = y - r; cy <= y + r; ++cy)
                for (int cx = x - r; cx <= x + r; ++cx)
                {
                    const float space2 = (float)((x - cx) * (x - cx) + (y - cy) * (y - cy));
                    if (space2 > r2) continue;
                    const float value = src[(size_t)reflect101(cy, h) * w + reflect101(cx, w)];
                    const float ad = fabsf(value - center);
                    const float wgt = expf(fmaf(space2, ss, (ad * ad) * sc));
                    sum1 = fmaf(wgt, value, sum1);
                    sum2 = sum2 + wgt;
                }
            dst[(size_t)y * w + x] = sum1 / sum2;
        }
}

/* image_process.cu:8-17 (without the out-of-bounds second test, §9 Q3) */
void kfo_truncate(float *d, int w, int h, float max_dist)
{
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i)
    {
        d[i] *= 0.001f;
        if (d[i] > max_dist) d[i] = 0.f;
    }
}

/* image_process.cu:29-43 + device_utils.cuh:22-27 */
void kfo_vertex_map(const float *d, const kfo_intr *k, float *vmap3)
{
    const int w = k->width, h = k->height;
    const float rfx = KFO_RCP(k->fx), rfy = KFO_RCP(k->fy);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const float z = d[(size_t)y * w + x];
            float *o = vmap3 + 3 * ((size_t)y * w + x);
            if (isnan(z)) { o[0] = o[1] = o[2] = 0.f; continue; }
            o[0] = rfx * (z * ((float)x - k->cx));
            o[1] = rfy * (z * ((float)y - k->cy));
            o[2] = z;
        }
}

/* image_process.cu:57-84; border pixels are the zeros left by Frame::reset
 * (types.hpp:53-62, §9 Q6); normalize = IEEE sqrtf and divisions
 * (device_types.hpp:253-257) so 0/0 = NaN marks invalid normals. */
void kfo_normal_map(const float *vmap3, int w, int h, float *nmap3)
{
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            float *o = nmap3 + 3 * ((size_t)y * w + x);
            if (x < 1 || x >= w - 1 || y < 1 || y >= h - 1) { o[0] = o[1] = o[2] = 0.f; continue; }
            const float *l = vmap3 + 3 * ((size_t)y * w + x - 1);
            const float *r = vmap3 + 3 * ((size_t)y * w + x + 1);
            const float *u = vmap3 + 3 * ((size_t)(y - 1) * w + x);
            const float *dn = vmap3 + 3 * ((size_t)(y + 1) * w + x);
            float nx = 0.f, ny = 0.f, nz = 0.f;
            if (!(l[2] == 0 || r[2] == 0 || u[2] == 0 || dn[2] == 0))
            {
                const float ax = l[0] - r[0], ay = l[1] - r[1], az = l[2] - r[2];
                const float bx = u[0] - dn[0], by = u[1] - dn[1], bz = u[2] - dn[2];
                nx = fmaf(ay, bz, -(az * by));
                ny = fmaf(az, bx, -(ax * bz));
                nz = fmaf(ax, by, -(ay * bx));
                if (nz > 0) { nx = -nx; ny = -ny; nz = -nz; }
            }
            const float t = sqrtf(dot3c(nx, ny, nz, nx, ny, nz));
            o[0] = nx / t; o[1] = ny / t; o[2] = nz / t;
        }
}

/* image_process.cu:95-125 (§9 Q18) */
void kfo_resize_maps(const float *vbig, const float *nbig, int bw, int bh, float *vsmall, float *nsmall)
{
    const int sw = bw >> 1, sh = bh >> 1;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < sh; ++y)
        for (int x = 0; x < sw; ++x)
        {
            float *vo = vsmall + 3 * ((size_t)y * sw + x);
            float *no = nsmall + 3 * ((size_t)y * sw + x);
            vo[0] = vo[1] = vo[2] = no[0] = no[1] = no[2] = 0.f;
            const size_t i00 = 3 * ((size_t)(2 * y) * bw + 2 * x), i01 = i00 + 3;
            const size_t i10 = i00 + 3 * (size_t)bw, i11 = i10 + 3;
            const float prod = ((vbig[i00] * vbig[i01]) * vbig[i10]) * vbig[i11];
            if (!isnan(prod))
                for (int c = 0; c < 3; ++c)
                {
                    vo[c] = (((vbig[i00 + c] + vbig[i01 + c]) + vbig[i10 + c]) + vbig[i11 + c]) * 0.25f;
                    no[c] = (((nbig[i00 + c] + nbig[i01 + c]) + nbig[i10 + c]) + nbig[i11 + c]) * 0.25f;
                }
        }
}

/* ======================================================================
 * ICP
 * ====================================================================== */

/* rigid_icp.cu:46-80 (findCoresp) + :81-95 (row).  returns 1 and fills row[7]. */
static inline int icp_row(const float *cur_v, const float *cur_n, const float *pre_v, const float *pre_n,
                          int w, int h, int x, int y, const kfo_intr *k, const float R[9], const float t[3],
                          float dist_thres, float sine_thres, float row[7])
{
    const size_t i = 3 * ((size_t)y * w + x);
    if (isnan(cur_n[i])) return 0;
    float vc[3];
    rot3(R, cur_v[i], cur_v[i + 1], cur_v[i + 2], vc);
    vc[0] += t[0]; vc[1] += t[1]; vc[2] += t[2];
    const float rz = KFO_RCP(vc[2]);
    const int px = f2i_rn(fmaf(rz * vc[0], k->fx, k->cx));
    const int py = f2i_rn(fmaf(rz * vc[1], k->fy, k->cy));
    if (!(vc[2] > 0 && px >= 0 && py >= 0 && px < w && py < h)) return 0;
    const size_t j = 3 * ((size_t)py * w + px);
    const float dx = vc[0] - pre_v[j], dy = vc[1] - pre_v[j + 1], dz = vc[2] - pre_v[j + 2];
    const float dist = sqrtf(dot3c(dx, dy, dz, dx, dy, dz));
    if (!(dist <= dist_thres)) return 0;
    float nc[3];
    rot3(R, cur_n[i], cur_n[i + 1], cur_n[i + 2], nc);
    const float npx = pre_n[j], npy = pre_n[j + 1], npz = pre_n[j + 2];
    const float sx = fmaf(nc[1], npz, -(nc[2] * npy));
    const float sy = fmaf(nc[2], npx, -(nc[0] * npz));
    const float sz = fmaf(nc[0], npy, -(nc[1] * npx));
    const float sine = sqrtf(dot3c(sx, sy, sz, sx, sy, sz));
    if (!(sine <= sine_thres)) return 0;
    /* row = [s x n, n, n.(d - s)], s = vcur, d = vpre, n = npre */
    row[0] = fmaf(vc[1], npz, -(vc[2] * npy));
    row[1] = fmaf(vc[2], npx, -(vc[0] * npz));
    row[2] = fmaf(vc[0], npy, -(vc[1] * npx));
    row[3] = npx; row[4] = npy; row[5] = npz;
    const float ex = pre_v[j] - vc[0], ey = pre_v[j + 1] - vc[1], ez = pre_v[j + 2] - vc[2];
    row[6] = fmaf(npz, ez, fmaf(npx, ex, npy * ey));
    return 1;
}

/* rigid_icp.cu:81-169: f32 products, per-32x32-tile f64 sum stored as f32,
 * per-thread f32 accumulate over tiles t, t+512, ..., f64 tree, stored f32 (§9 Q10).
 * compat_rows=1 reproduces the truncated grid (§9 Q7); 0 covers every pixel. */
void kfo_icp_accumulate(const float *cur_v, const float *cur_n, const float *pre_v, const float *pre_n,
                        const kfo_intr *k, const float pose12[12], float dist_thres, float sine_thres,
                        int compat_rows, double out27[27], int64_t *n_corresp)
{
    const int w = k->width, h = k->height;
    float R[9];
    pose_R9(pose12, R);
    const float t[3] = {pose12[3], pose12[7], pose12[11]};
    const int tx = compat_rows ? w / 32 : (w + 31) / 32;
    const int ty = compat_rows ? h / 32 : (h + 31) / 32;
    const int ntiles = tx * ty;
    float *tile = (float *)calloc((size_t)(ntiles > 0 ? ntiles : 1) * 27, sizeof(float));
    int64_t count = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : count)
    for (int tl = 0; tl < ntiles; ++tl)
    {
        const int bx = tl % tx, by = tl / tx;
        double acc[27];
        for (int i = 0; i < 27; ++i) acc[i] = 0.0;
        for (int yy = 0; yy < 32; ++yy)
            for (int xx = 0; xx < 32; ++xx)
            {
                const int x = bx * 32 + xx, y = by * 32 + yy;
                if (x >= w || y >= h) continue;
                float row[7];
                if (!icp_row(cur_v, cur_n, pre_v, pre_n, w, h, x, y, k, R, t, dist_thres, sine_thres, row)) continue;
                ++count;
                int s = 0;
                for (int i = 0; i < 6; ++i)
                    for (int j = i; j < 7; ++j)
                        acc[s++] += (double)(row[i] * row[j]);
            }
        for (int i = 0; i < 27; ++i) tile[(size_t)i * ntiles + tl] = (float)acc[i];
    }
    for (int i = 0; i < 27; ++i)
    {
        float part[512];
        for (int p = 0; p < 512; ++p) part[p] = 0.f;
        for (int tl = 0; tl < ntiles; ++tl) part[tl & 511] += tile[(size_t)i * ntiles + tl];
        double s = 0.0;
        for (int p = 0; p < 512; ++p) s += (double)part[p];
        out27[i] = (double)(float)s;
    }
    free(tile);
    if (n_corresp) *n_corresp = count;
}

/* icp_registration.cpp:30-39 + rigid_icp.cu:156-165: unpack, determinant guard
 * (LU, double), solve.  cv::solve(DECOMP_SVD) is replaced by Cholesky in double
 * with an LU fallback (north star; equivalent for the non-singular case, §9 Q13). */
int kfo_icp_solve(const double in27[27], double x6[6])
{
    double A[6][6], b[6];
    int s = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j)
        {
            const double v = in27[s++];
            if (j == 6) b[i] = v;
            else A[i][j] = A[j][i] = v;
        }
    /* determinant by LU with partial pivoting */
    double M[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = A[i][j]; M[i][6] = b[i]; }
    double det = 1.0;
    int singular = 0;
    for (int c = 0; c < 6; ++c)
    {
        int p = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        if (M[p][c] == 0.0 || isnan(M[p][c])) { det = isnan(M[p][c]) ? NAN : 0.0; singular = 1; break; }
        if (p != c) { for (int j = 0; j < 7; ++j) { double tmp = M[c][j]; M[c][j] = M[p][j]; M[p][j] = tmp; } det = -det; }
        det *= M[c][c];
        for (int r = c + 1; r < 6; ++r)
        {
            const double f = M[r][c] / M[c][c];
            for (int j = c; j < 7; ++j) M[r][j] -= f * M[c][j];
        }
    }
    if (singular || fabs(det) < 1e-15 || isnan(det)) return 1;
    /* Cholesky A = L L^T */
    double L[6][6];
    int ok = 1;
    memset(L, 0, sizeof(L));
    for (int i = 0; i < 6 && ok; ++i)
        for (int j = 0; j <= i; ++j)
        {
            double sum = A[i][j];
            for (int q = 0; q < j; ++q) sum -= L[i][q] * L[j][q];
            if (i == j) { if (!(sum > 0.0)) { ok = 0; break; } L[i][i] = sqrt(sum); }
            else L[i][j] = sum / L[j][j];
        }
    if (ok)
    {
        double yv[6];
        for (int i = 0; i < 6; ++i) { double sum = b[i]; for (int q = 0; q < i; ++q) sum -= L[i][q] * yv[q]; yv[i] = sum / L[i][i]; }
        for (int i = 5; i >= 0; --i) { double sum = yv[i]; for (int q = i + 1; q < 6; ++q) sum -= L[q][i] * x6[q]; x6[i] = sum / L[i][i]; }
    }
    else
    {
        for (int i = 5; i >= 0; --i) { double sum = M[i][6]; for (int q = i + 1; q < 6; ++q) sum -= M[i][q] * x6[q]; x6[i] = sum / M[i][i]; }
    }
    return 0;
}

/* ======================================================================
 * Pose algebra (restated OpenCV core)
 * ====================================================================== */
void kfo_pose_identity(float p[12])
{
    memset(p, 0, 12 * sizeof(float));
    p[0] = p[5] = p[10] = 1.f;
}
/* cv::Affine3f operator* = 4x4 float matrix product */
void kfo_pose_mul(const float a[12], const float b[12], float out[12])
{
    float r[12];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j)
        {
            float s = 0.f;
            for (int q = 0; q < 3; ++q) s += a[4 * i + q] * b[4 * q + j];
            if (j == 3) s += a[4 * i + 3];
            r[4 * i + j] = s;
        }
    memcpy(out, r, sizeof(r));
}
/* cv::Affine3f::inv(): general inverse of [A|t], evaluated in double */
void kfo_pose_inv(const float a[12], float out[12])
{
    const double m00 = a[0], m01 = a[1], m02 = a[2], m10 = a[4], m11 = a[5], m12 = a[6], m20 = a[8], m21 = a[9], m22 = a[10];
    const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    const double id = 1.0 / det;
    double inv[9];
    inv[0] = c00 * id; inv[1] = (m02 * m21 - m01 * m22) * id; inv[2] = (m01 * m12 - m02 * m11) * id;
    inv[3] = c01 * id; inv[4] = (m00 * m22 - m02 * m20) * id; inv[5] = (m02 * m10 - m00 * m12) * id;
    inv[6] = c02 * id; inv[7] = (m01 * m20 - m00 * m21) * id; inv[8] = (m00 * m11 - m01 * m10) * id;
    const double tx = a[3], ty = a[7], tz = a[11];
    for (int i = 0; i < 3; ++i)
    {
        out[4 * i + 0] = (float)inv[3 * i + 0];
        out[4 * i + 1] = (float)inv[3 * i + 1];
        out[4 * i + 2] = (float)inv[3 * i + 2];
        out[4 * i + 3] = (float)(-(inv[3 * i] * tx + inv[3 * i + 1] * ty + inv[3 * i + 2] * tz));
    }
}
/* cam2vol.rotation().inv(DECOMP_SVD) (tsdf_volume.cpp:61) */
void kfo_rot_inv(const float a[12], float rinv9[9])
{
    float tmp[12], in[12];
    memcpy(in, a, sizeof(in));
    in[3] = in[7] = in[11] = 0.f;
    kfo_pose_inv(in, tmp);
    pose_R9(tmp, rinv9);
}
/* cv::Affine3f(rvec, t) rotation: Rodrigues in double on float rvec */
void kfo_rodrigues(const float rvec[3], float R9[9])
{
    const double rx = rvec[0], ry = rvec[1], rz = rvec[2];
    const double theta = sqrt(rx * rx + ry * ry + rz * rz);
    if (theta < DBL_EPSILON)
    {
        for (int i = 0; i < 9; ++i) R9[i] = (i % 4 == 0) ? 1.f : 0.f;
        return;
    }
    const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
    const double x = rx * it, y = ry * it, z = rz * it;
    R9[0] = (float)(c + c1 * x * x);     R9[1] = (float)(c1 * x * y - s * z); R9[2] = (float)(c1 * x * z + s * y);
    R9[3] = (float)(c1 * x * y + s * z); R9[4] = (float)(c + c1 * y * y);     R9[5] = (float)(c1 * y * z - s * x);
    R9[6] = (float)(c1 * x * z - s * y); R9[7] = (float)(c1 * y * z + s * x); R9[8] = (float)(c + c1 * z * z);
}
/* icp_registration.cpp:41-42: Tinc = Affine3f(rvec=x[0..2], t=x[3..5]); pose = pose * Tinc */
void kfo_pose_apply_increment(float pose12[12], const double x6[6])
{
    const float rv[3] = {(float)x6[0], (float)x6[1], (float)x6[2]};
    float R9[9], inc[12];
    kfo_rodrigues(rv, R9);
    for (int i = 0; i < 3; ++i)
    {
        inc[4 * i + 0] = R9[3 * i + 0]; inc[4 * i + 1] = R9[3 * i + 1]; inc[4 * i + 2] = R9[3 * i + 2];
        inc[4 * i + 3] = (float)x6[3 + i];
    }
    kfo_pose_mul(pose12, inc, pose12);
}

/* ======================================================================
 * TSDF integrate -- tsdf_volume.cu:41-99 (§9 Q15), colour dropped (§9 Q16)
 * z_begin/z_end select a slab [z_begin, z_end) of planes to WRITE; the running
 * sum is always replayed from z=1 so slab results equal the full sweep.
 * ====================================================================== */
void kfo_integrate(int16_t *vol, const kfo_volume_desc *vd, const float vol2cam[12],
                   const float *depth_m, const kfo_intr *k, int z_begin, int z_end, int64_t *n_updated)
{
    const int X = vd->dims[0], Y = vd->dims[1], Z = vd->dims[2];
    const int w = k->width, h = k->height;
    float R[9];
    pose_R9(vol2cam, R);
    const float tx = vol2cam[3], ty = vol2cam[7], tz = vol2cam[11];
    const float vsx = vd->voxel_size[0], vsy = vd->voxel_size[1], vsz = vd->voxel_size[2];
    const float trunc = vd->trunc_dist;
    const float rtrunc = KFO_RCP(trunc);
    const float rfx = KFO_RCP(k->fx), rfy = KFO_RCP(k->fy);
    const size_t plane = (size_t)X * Y;
    int64_t upd = 0;
    if (z_begin < 1) z_begin = 1;
    if (z_end > Z) z_end = Z;
#pragma omp parallel for schedule(static) reduction(+ : upd)
    for (int y = 0; y < Y; ++y)
        for (int x = 0; x < X; ++x)
        {
            const float px = (float)x * vsx, py = (float)y * vsy, pz = 0.f * vsz;
            float r3[3];
            rot3(R, px, py, pz, r3);
            float vcx = r3[0] + tx, vcy = r3[1] + ty, vcz = r3[2] + tz;
            int16_t *vp = vol + 2 * ((size_t)x + (size_t)y * X);
            for (int z = 1; z < z_end; ++z)
            {
                vp += 2 * plane;
                /* vc += zstep, contracted to fma(voxel_size.x, R[:,2], vc) */
                vcx = fmaf(vsx, R[2], vcx);
                vcy = fmaf(vsx, R[5], vcy);
                vcz = fmaf(vsx, R[8], vcz);
                if (z < z_begin) continue;
                if (vcz <= 0) continue;
                const float rz = KFO_RCP(vcz);
                const int u = f2i_rn(fmaf(rz * vcx, k->fx, k->cx));
                const int v = f2i_rn(fmaf(rz * vcy, k->fy, k->cy));
                if (u < 0 || u >= w || v < 0 || v >= h) continue;
                const float depth = depth_m[(size_t)v * w + u];
                if (depth <= 0) continue;
                const float lx = rfx * ((float)u - k->cx);
                const float ly = rfy * ((float)v - k->cy);
                const float lambda = sqrtf(fmaf(lx, lx, ly * ly) + 1.0f);
                const float nrm = sqrtf(dot3c(vcx, vcy, vcz, vcx, vcy, vcz));
                const float nsdf = fmaf(KFO_RCP(lambda), nrm, -depth); /* = -sdf */
                if (!(nsdf <= trunc)) continue;
                const float tsdf = fminf(1.f, rtrunc * (-nsdf));
                const float pre = (float)vp[0] * KFO_DIVSHORTMAX;
                const int pw = vp[1];
                int nw = pw + 1;
                const float nt = KFO_RCP((float)nw) * fmaf(pre, (float)pw, tsdf);
                if (nw > vd->max_weight) nw = vd->max_weight;
                int q = f2i_rz(nt * (float)KFO_SHORTMAX);
                if (q > KFO_SHORTMAX) q = KFO_SHORTMAX;
                if (q < -KFO_SHORTMAX) q = -KFO_SHORTMAX;
                vp[0] = (int16_t)q;
                vp[1] = (int16_t)nw;
                ++upd;
            }
        }
    if (n_updated) *n_updated = upd;
}

/* ======================================================================
 * Raycast -- tsdf_volume.cu:120-273 (§9 Q17)
 * ====================================================================== */
typedef struct
{
    const int16_t *vol;
    int X, Y, Z;
    size_t plane;
    float vsinv[3], gd[3];
    int zs0, zs1; /* stored planes [zs0, zs1): vol points at plane zs0 (z-slab shards) */
    int zo0, zo1; /* owned planes */
} rc_ctx;

static inline float vox_tsdf(const rc_ctx *c, int x, int y, int z)
{
    if (z < c->zs0) z = c->zs0;
    if (z >= c->zs1) z = c->zs1 - 1;
    return (float)c->vol[2 * ((size_t)x + (size_t)y * c->X + (size_t)(z - c->zs0) * c->plane)] * KFO_DIVSHORTMAX;
}
/* tsdf_volume.cu:178-191 */
static inline float voxel2tsdf(const rc_ctx *c, float px, float py, float pz)
{
    const int x = f2i_rn(px * c->vsinv[0]);
    const int y = f2i_rn(py * c->vsinv[1]);
    const int z = f2i_rn(pz * c->vsinv[2]);
    if (x >= c->X - 1 || y >= c->Y - 1 || z >= c->Z - 1 || x < 1 || y < 1 || z < 1) return NAN;
    return vox_tsdf(c, x, y, z);
}
/* slab variant: NaN outside the stored planes; *own = the sample's voxel plane is owned by the slab */
static inline float voxel2tsdf_slab(const rc_ctx *c, float px, float py, float pz, int *own)
{
    const int x = f2i_rn(px * c->vsinv[0]);
    const int y = f2i_rn(py * c->vsinv[1]);
    const int z = f2i_rn(pz * c->vsinv[2]);
    *own = 0;
    if (x >= c->X - 1 || y >= c->Y - 1 || z >= c->Z - 1 || x < 1 || y < 1 || z < 1) return NAN;
    if (z < c->zs0 || z >= c->zs1) return NAN;
    *own = z >= c->zo0 && z < c->zo1;
    return vox_tsdf(c, x, y, z);
}
/* tsdf_volume.cu:137-161 */
static inline float interpolate(const rc_ctx *c, float fx, float fy, float fz)
{
    const int gx = f2i_rd(fx), gy = f2i_rd(fy), gz = f2i_rd(fz);
    if (gx < 0 || gx >= c->X - 1 || gy < 0 || gy >= c->Y - 1 || gz < 0 || gz >= c->Z - 1) return NAN;
    const float a = fx - (float)gx, b = fy - (float)gy, cc = fz - (float)gz;
    const float a1 = 1.f - a, b1 = 1.f - b, c1 = 1.f - cc;
    float t = 0.f;
    t = fmaf((vox_tsdf(c, gx, gy, gz) * a1) * b1, c1, t);
    t = fmaf((vox_tsdf(c, gx, gy, gz + 1) * a1) * b1, cc, t);
    t = fmaf((vox_tsdf(c, gx, gy + 1, gz) * a1) * b, c1, t);
    t = fmaf((vox_tsdf(c, gx, gy + 1, gz + 1) * a1) * b, cc, t);
    t = fmaf((vox_tsdf(c, gx + 1, gy, gz) * a) * b1, c1, t);
    t = fmaf((vox_tsdf(c, gx + 1, gy, gz + 1) * a) * b1, cc, t);
    t = fmaf((vox_tsdf(c, gx + 1, gy + 1, gz) * a) * b, c1, t);
    t = fmaf((vox_tsdf(c, gx + 1, gy + 1, gz + 1) * a) * b, cc, t);
    return t;
}

/* Slab semantics (test model of kfb_raycast on a z-slab context + kfb_composite_mask): `vol` holds planes
 * [zs0, zs1) only; a sample outside them is NaN; a (cur, next) pair is evaluated only if next's voxel plane
 * is in [zo0, zo1); key[pixel] = ray_len of the first terminal event (hit or back-face stop), +inf if none.
 * With zs = [0, Z) and zo = [0, Z) this is the reference's raycast. */
void kfo_raycast_slab(const int16_t *vol, const kfo_volume_desc *vd, const float cam2vol[12], const float rinv9[9],
                      const kfo_intr *k, float *vmap3, float *nmap3, float *key, int zs0, int zs1, int zo0, int zo1,
                      int compat_ts_sign);

void kfo_raycast(const int16_t *vol, const kfo_volume_desc *vd, const float cam2vol[12], const float rinv9[9],
                 const kfo_intr *k, float *vmap3, float *nmap3, int64_t *n_steps, int compat_ts_sign)
{
    rc_ctx c;
    c.vol = vol;
    c.zs0 = 0; c.zs1 = vd->dims[2]; c.zo0 = 0; c.zo1 = vd->dims[2];
    c.X = vd->dims[0]; c.Y = vd->dims[1]; c.Z = vd->dims[2];
    c.plane = (size_t)c.X * c.Y;
    for (int i = 0; i < 3; ++i) { c.vsinv[i] = 1.f / vd->voxel_size[i]; c.gd[i] = vd->voxel_size[i] * 0.5f; }
    const float step_len = vd->voxel_size[0];
    float R[9];
    pose_R9(cam2vol, R);
    const float ox = cam2vol[3], oy = cam2vol[7], oz = cam2vol[11];
    const int w = k->width, h = k->height;
    const float rfx = KFO_RCP(k->fx), rfy = KFO_RCP(k->fy);
    const float rgx = KFO_RCP(c.gd[0]), rgy = KFO_RCP(c.gd[1]), rgz = KFO_RCP(c.gd[2]);
    int64_t steps = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : steps)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            float *vo = vmap3 + 3 * ((size_t)y * w + x);
            float *no = nmap3 + 3 * ((size_t)y * w + x);
            vo[0] = vo[1] = vo[2] = no[0] = no[1] = no[2] = 0.f; /* pframe->reset(), kinectfusion.cpp:112 */
            const float px = rfx * ((float)x - k->cx), py = rfy * ((float)y - k->cy);
            float dx = fmaf(px, R[0], py * R[1]) + R[2];
            float dy = fmaf(px, R[3], py * R[4]) + R[5];
            float dz = fmaf(px, R[6], py * R[7]) + R[8];
            {
                const float rt = KFO_RCP(sqrtf(dot3c(dx, dy, dz, dx, dy, dz)));
                dx = rt * dx; dy = rt * dy; dz = rt * dz;
            }
            /* intersect(), tsdf_volume.cu:120-136 */
            const float ix = 1.f / dx, iy = 1.f / dy, iz = 1.f / dz;
            const float bx = ix * (0.f - ox), by = iy * (0.f - oy), bz = iz * (0.f - oz);
            const float tx_ = ix * (vd->range[0] - ox), ty_ = iy * (vd->range[1] - oy), tz_ = iz * (vd->range[2] - oz);
            const float mnx = fminf(tx_, bx), mny = fminf(ty_, by), mnz = fminf(tz_, bz);
            const float mxx = fmaxf(tx_, bx), mxy = fmaxf(ty_, by), mxz = fmaxf(tz_, bz);
            const float tnear = fmaxf(fmaxf(mnx, mny), fmaxf(mnx, mnz));
            const float tfar = fminf(fminf(mxx, mxy), fminf(mxx, mxz));
            float ray_len = fmaxf(tnear, 0.f);
            if (ray_len >= tfar) continue;
            ray_len += step_len;
            float nx_ = fmaf(dx, ray_len, ox), ny_ = fmaf(dy, ray_len, oy), nz_ = fmaf(dz, ray_len, oz);
            float tnext = voxel2tsdf(&c, nx_, ny_, nz_);
            for (; ray_len < tfar; ray_len += step_len)
            {
                ++steps;
                nx_ = fmaf(dx, vd->voxel_size[0], nx_);
                ny_ = fmaf(dy, vd->voxel_size[1], ny_);
                nz_ = fmaf(dz, vd->voxel_size[2], nz_);
                const float tcur = tnext;
                tnext = voxel2tsdf(&c, nx_, ny_, nz_);
                if (isnan(tnext)) continue;
                if (tcur < 0.f && tnext > 0.f) break;
                if (tcur > 0.f && tnext < 0.f)
                {
                    const float q = KFO_RCP(tcur - tnext);
                    const float num = tcur * vd->voxel_size[0];
                    const float Ts = compat_ts_sign ? fmaf(q, -num, ray_len) : fmaf(q, num, ray_len);
                    const float vx = fmaf(dx, Ts, ox), vy = fmaf(dy, Ts, oy), vz = fmaf(dz, Ts, oz);
                    /* compute_normal(), tsdf_volume.cu:192-209 */
                    const float Fx1 = interpolate(&c, (vx + c.gd[0]) * c.vsinv[0], vy * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fx2 = interpolate(&c, (vx - c.gd[0]) * c.vsinv[0], vy * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fy1 = interpolate(&c, vx * c.vsinv[0], (vy + c.gd[1]) * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fy2 = interpolate(&c, vx * c.vsinv[0], (vy - c.gd[1]) * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fz1 = interpolate(&c, vx * c.vsinv[0], vy * c.vsinv[1], (vz + c.gd[2]) * c.vsinv[2]);
                    const float Fz2 = interpolate(&c, vx * c.vsinv[0], vy * c.vsinv[1], (vz - c.gd[2]) * c.vsinv[2]);
                    float gx = rgx * (Fx1 - Fx2), gy = rgy * (Fy1 - Fy2), gz = rgz * (Fz1 - Fz2);
                    const float rn = KFO_RCP(sqrtf(dot3c(gx, gy, gz, gx, gy, gz)));
                    gx = rn * gx; gy = rn * gy; gz = rn * gz;
                    if (!isnan((gx * gy) * gz))
                    {
                        rot3(rinv9, gx, gy, gz, no);
                        rot3(rinv9, vx - ox, vy - oy, vz - oz, vo);
                        break;
                    }
                }
            }
        }
    if (n_steps) *n_steps = steps;
}

void kfo_raycast_slab(const int16_t *vol, const kfo_volume_desc *vd, const float cam2vol[12], const float rinv9[9],
                      const kfo_intr *k, float *vmap3, float *nmap3, float *key, int zs0, int zs1, int zo0, int zo1,
                      int compat_ts_sign)
{
    rc_ctx c;
    c.vol = vol;
    c.zs0 = zs0; c.zs1 = zs1; c.zo0 = zo0; c.zo1 = zo1;
    c.X = vd->dims[0]; c.Y = vd->dims[1]; c.Z = vd->dims[2];
    c.plane = (size_t)c.X * c.Y;
    for (int i = 0; i < 3; ++i) { c.vsinv[i] = 1.f / vd->voxel_size[i]; c.gd[i] = vd->voxel_size[i] * 0.5f; }
    const float step_len = vd->voxel_size[0];
    float R[9];
    pose_R9(cam2vol, R);
    const float ox = cam2vol[3], oy = cam2vol[7], oz = cam2vol[11];
    const int w = k->width, h = k->height;
    const float rfx = KFO_RCP(k->fx), rfy = KFO_RCP(k->fy);
    const float rgx = KFO_RCP(c.gd[0]), rgy = KFO_RCP(c.gd[1]), rgz = KFO_RCP(c.gd[2]);
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            float *vo = vmap3 + 3 * ((size_t)y * w + x);
            float *no = nmap3 + 3 * ((size_t)y * w + x);
            vo[0] = vo[1] = vo[2] = no[0] = no[1] = no[2] = 0.f;
            key[(size_t)y * w + x] = INFINITY;
            const float px = rfx * ((float)x - k->cx), py = rfy * ((float)y - k->cy);
            float dx = fmaf(px, R[0], py * R[1]) + R[2];
            float dy = fmaf(px, R[3], py * R[4]) + R[5];
            float dz = fmaf(px, R[6], py * R[7]) + R[8];
            {
                const float rt = KFO_RCP(sqrtf(dot3c(dx, dy, dz, dx, dy, dz)));
                dx = rt * dx; dy = rt * dy; dz = rt * dz;
            }
            const float ix = 1.f / dx, iy = 1.f / dy, iz = 1.f / dz;
            const float bx = ix * (0.f - ox), by = iy * (0.f - oy), bz = iz * (0.f - oz);
            const float tx_ = ix * (vd->range[0] - ox), ty_ = iy * (vd->range[1] - oy), tz_ = iz * (vd->range[2] - oz);
            const float mnx = fminf(tx_, bx), mny = fminf(ty_, by), mnz = fminf(tz_, bz);
            const float mxx = fmaxf(tx_, bx), mxy = fmaxf(ty_, by), mxz = fmaxf(tz_, bz);
            const float tnear = fmaxf(fmaxf(mnx, mny), fmaxf(mnx, mnz));
            const float tfar = fminf(fminf(mxx, mxy), fminf(mxx, mxz));
            float ray_len = fmaxf(tnear, 0.f);
            if (ray_len >= tfar) continue;
            ray_len += step_len;
            float nx_ = fmaf(dx, ray_len, ox), ny_ = fmaf(dy, ray_len, oy), nz_ = fmaf(dz, ray_len, oz);
            int own;
            float tnext = voxel2tsdf_slab(&c, nx_, ny_, nz_, &own);
            for (; ray_len < tfar; ray_len += step_len)
            {
                nx_ = fmaf(dx, vd->voxel_size[0], nx_);
                ny_ = fmaf(dy, vd->voxel_size[1], ny_);
                nz_ = fmaf(dz, vd->voxel_size[2], nz_);
                const float tcur = tnext;
                tnext = voxel2tsdf_slab(&c, nx_, ny_, nz_, &own);
                if (isnan(tnext) || !own) continue;
                if (tcur < 0.f && tnext > 0.f) { key[(size_t)y * w + x] = ray_len; break; }
                if (tcur > 0.f && tnext < 0.f)
                {
                    const float q = KFO_RCP(tcur - tnext);
                    const float num = tcur * vd->voxel_size[0];
                    const float Ts = compat_ts_sign ? fmaf(q, -num, ray_len) : fmaf(q, num, ray_len);
                    const float vx = fmaf(dx, Ts, ox), vy = fmaf(dy, Ts, oy), vz = fmaf(dz, Ts, oz);
                    const float Fx1 = interpolate(&c, (vx + c.gd[0]) * c.vsinv[0], vy * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fx2 = interpolate(&c, (vx - c.gd[0]) * c.vsinv[0], vy * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fy1 = interpolate(&c, vx * c.vsinv[0], (vy + c.gd[1]) * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fy2 = interpolate(&c, vx * c.vsinv[0], (vy - c.gd[1]) * c.vsinv[1], vz * c.vsinv[2]);
                    const float Fz1 = interpolate(&c, vx * c.vsinv[0], vy * c.vsinv[1], (vz + c.gd[2]) * c.vsinv[2]);
                    const float Fz2 = interpolate(&c, vx * c.vsinv[0], vy * c.vsinv[1], (vz - c.gd[2]) * c.vsinv[2]);
                    float gx = rgx * (Fx1 - Fx2), gy = rgy * (Fy1 - Fy2), gz = rgz * (Fz1 - Fz2);
                    const float rn = KFO_RCP(sqrtf(dot3c(gx, gy, gz, gx, gy, gz)));
                    gx = rn * gx; gy = rn * gy; gz = rn * gz;
                    if (!isnan((gx * gy) * gz))
                    {
                        rot3(rinv9, gx, gy, gz, no);
                        rot3(rinv9, vx - ox, vy - oy, vz - oz, vo);
                        key[(size_t)y * w + x] = ray_len;
                        break;
                    }
                }
            }
        }
}

/* ======================================================================
 * Point-cloud extraction -- tsdf_volume.cu:328-419 (§9 Q21).  Output order is
 * nondeterministic in the reference (atomics); here it is x-fastest scan order,
 * compare as sorted sets.
 * ====================================================================== */
int64_t kfo_extract_points(const int16_t *vol, const kfo_volume_desc *vd, const float volpose[12],
                           float *points3, int64_t cap)
{
    const int X = vd->dims[0], Y = vd->dims[1], Z = vd->dims[2];
    const size_t plane = (size_t)X * Y;
    float R[9];
    pose_R9(volpose, R);
    const float t[3] = {volpose[3], volpose[7], volpose[11]};
    int64_t n = 0;
    for (int z = 0; z < Z - 1; ++z)
        for (int y = 0; y < Y; ++y)
            for (int x = 0; x < X; ++x)
            {
                const size_t i = (size_t)x + (size_t)y * X + (size_t)z * plane;
                const int W = vol[2 * i + 1];
                const float F = (float)vol[2 * i] * KFO_DIVSHORTMAX;
                if (W == 0 || F == 1.f) continue;
                const float Vx = ((float)x + 0.5f) * vd->voxel_size[0];
                const float Vy = ((float)y + 0.5f) * vd->voxel_size[1];
                const float Vz = ((float)z + 0.5f) * vd->voxel_size[2];
                for (int axis = 0; axis < 3; ++axis)
                {
                    size_t j;
                    if (axis == 0) { if (x + 1 >= X) continue; j = i + 1; }
                    else if (axis == 1) { if (y + 1 >= Y) continue; j = i + X; }
                    else j = i + plane;
                    const int Wn = vol[2 * j + 1];
                    const float Fn = (float)vol[2 * j] * KFO_DIVSHORTMAX;
                    if (Wn == 0 || Fn == 1.f) continue;
                    if (!((F > 0 && Fn < 0) || (F < 0 && Fn > 0))) continue;
                    const float Vv[3] = {Vx, Vy, Vz};
                    const float V = Vv[axis];
                    const float Vn = V + vd->voxel_size[axis];
                    const float d_inv = 1.f / (fabsf(F) + fabsf(Fn));
                    /* (V*|Fn| + Vn*|F|) * d_inv as compiled: V*|Fn| rounded, |F|*Vn fused onto it */
                    const float pi = fmaf(fabsf(F), Vn, V * fabsf(Fn)) * d_inv;
                    if (n < cap)
                    {
                        /* aff.R * p + aff.t as compiled (loop-invariant products hoisted and rounded):
                         *  x: fma(Vz,R2, fma(px,R0, Vy*R1));  y: fma(Vz,R2, fma(py,R1, Vx*R0));
                         *  z: fma(pz,R2, Vy*R1 + Vx*R0) */
                        for (int c = 0; c < 3; ++c)
                        {
                            const float r0 = R[3 * c], r1 = R[3 * c + 1], r2 = R[3 * c + 2];
                            float o;
                            if (axis == 0) o = fmaf(Vz, r2, fmaf(pi, r0, Vy * r1));
                            else if (axis == 1) o = fmaf(Vz, r2, fmaf(pi, r1, Vx * r0));
                            else o = fmaf(pi, r2, Vy * r1 + Vx * r0);
                            points3[3 * n + c] = o + t[c];
                        }
                    }
                    ++n;
                }
            }
    return n < cap ? n : cap;
}

/* image_process.cu:137-147 */
void kfo_render_normals(const float *nmap3, int w, int h, uint8_t *bgr)
{
    const size_t n = (size_t)w * h * 3;
    for (size_t i = 0; i < n; ++i)
    {
        int v = f2i_rz(fabsf(nmap3[i]) * 255.f);
        bgr[i] = (uint8_t)(v & 0xff);
    }
}
/* image_process.cu:159-211 (double-precision literals `0.9`, `0.5`, pow() kept) */
void kfo_render_phong(const float *vmap3, const float *nmap3, int w, int h, const float eye[3], uint8_t *bgr)
{
    const size_t npx = (size_t)w * h;
    for (size_t i = 0; i < npx; ++i)
    {
        const float *v = vmap3 + 3 * i, *n = nmap3 + 3 * i;
        uint8_t *o = bgr + 3 * i;
        if (n[0] == 0 && n[1] == 0 && n[2] == 0) continue;
        if (v[0] == 0 && v[1] == 0 && v[2] == 0) continue;
        float ex = eye[0] - v[0], ey = eye[1] - v[1], ez = eye[2] - v[2];
        float lx = 500.f - v[0], ly = 500.f - v[1], lz = -500.f - v[2];
        float r = KFO_RCP(sqrtf(dot3c(ex, ey, ez, ex, ey, ez)));
        ex *= r; ey *= r; ez *= r;
        r = KFO_RCP(sqrtf(dot3c(lx, ly, lz, lx, ly, lz)));
        lx *= r; ly *= r; lz *= r;
        float lc = dot3c(n[0], n[1], n[2], lx, ly, lz);
        if (lc <= 0) lc = -lc;
        const float light_intensity = (float)0.9;
        float coef = light_intensity * lc;
        const float dfx = 0.3843f * coef, dfy = 0.4745f * coef, dfz = 0.580f * coef;
        float hx = lx + ex, hy = ly + ey, hz = lz + ez;
        r = KFO_RCP(sqrtf(dot3c(hx, hy, hz, hx, hy, hz)));
        hx *= r; hy *= r; hz *= r;
        float hc = dot3c(n[0], n[1], n[2], hx, hy, hz);
        if (hc < 0) hc = -hc;
        coef = light_intensity * powf(hc, 10.f);
        const float sp = (float)(0.5 * (double)coef);
        const float kx = fminf(1.f, 0.1f + dfx + sp), ky = fminf(1.f, 0.1f + dfy + sp), kz = fminf(1.f, 0.1f + dfz + sp);
        o[0] = (uint8_t)f2i_rz(kx * 255.f);
        o[1] = (uint8_t)f2i_rz(ky * 255.f);
        o[2] = (uint8_t)f2i_rz(kz * 255.f);
    }
}

/* ======================================================================
 * Synthetic scene and trajectory -- double precision
 * ====================================================================== */
void kfo_trajectory_pose(int k, int period, float pose12[12])
{
    const double th = 2.0 * M_PI * (double)k / (double)period;
    const double yaw = (6.0 * M_PI / 180.0) * sin(th), pitch = (3.0 * M_PI / 180.0) * sin(2.0 * th);
    const double cy = cos(yaw), sy = sin(yaw), cp = cos(pitch), sp = sin(pitch);
    /* R = Ry(yaw) * Rx(pitch) */
    const double R[9] = {cy, sy * sp, sy * cp, 0.0, cp, -sp, -sy, cy * sp, cy * cp};
    const double t[3] = {0.10 * sin(th), 0.05 * sin(2.0 * th), 0.08 * (1.0 - cos(th))};
    for (int i = 0; i < 3; ++i)
    {
        pose12[4 * i + 0] = (float)R[3 * i]; pose12[4 * i + 1] = (float)R[3 * i + 1]; pose12[4 * i + 2] = (float)R[3 * i + 2];
        pose12[4 * i + 3] = (float)t[i];
    }
}

static inline void hit_min(double t, double *best) { if (t > 1e-9 && t < *best) *best = t; }

/* Inside-out room (walls x=+-1.3, y=+-1.1, back wall z=3.1), sphere c=(0.35,0.15,1.9) r=0.35,
 * box [-0.8,-0.2]x[0.3,1.1]x[1.6,2.2].  Output: z-depth in mm, rounded to uint16, stored f32
 * (the reference's ingest type, depth_sensor.cpp:192). */
void kfo_render_depth_mm(const float cam2world[12], const kfo_intr *k, float *depth_mm)
{
    const int w = k->width, h = k->height;
    double R[9], o[3];
    for (int i = 0; i < 3; ++i)
    {
        R[3 * i] = cam2world[4 * i]; R[3 * i + 1] = cam2world[4 * i + 1]; R[3 * i + 2] = cam2world[4 * i + 2];
        o[i] = cam2world[4 * i + 3];
    }
    const double bmin[3] = {-0.8, 0.3, 1.6}, bmax[3] = {-0.2, 1.1, 2.2};
    const double sc[3] = {0.35, 0.15, 1.9}, sr = 0.35;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x)
        {
            const double dc[3] = {((double)x - k->cx) / k->fx, ((double)y - k->cy) / k->fy, 1.0};
            double d[3];
            for (int i = 0; i < 3; ++i) d[i] = R[3 * i] * dc[0] + R[3 * i + 1] * dc[1] + R[3 * i + 2] * dc[2];
            double best = 1e30;
            /* room planes (seen from inside) */
            if (d[0] > 0) hit_min((1.3 - o[0]) / d[0], &best);
            if (d[0] < 0) hit_min((-1.3 - o[0]) / d[0], &best);
            if (d[1] > 0) hit_min((1.1 - o[1]) / d[1], &best);
            if (d[1] < 0) hit_min((-1.1 - o[1]) / d[1], &best);
            if (d[2] > 0) hit_min((3.1 - o[2]) / d[2], &best);
            /* sphere */
            {
                const double oc[3] = {o[0] - sc[0], o[1] - sc[1], o[2] - sc[2]};
                const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                const double b = 2.0 * (oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2]);
                const double cc = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sr * sr;
                const double disc = b * b - 4.0 * a * cc;
                if (disc >= 0) hit_min((-b - sqrt(disc)) / (2.0 * a), &best);
            }
            /* box (slab test) */
            {
                double t0 = -1e30, t1 = 1e30;
                int ok = 1;
                for (int i = 0; i < 3; ++i)
                {
                    if (fabs(d[i]) < 1e-12) { if (o[i] < bmin[i] || o[i] > bmax[i]) ok = 0; continue; }
                    double a = (bmin[i] - o[i]) / d[i], b = (bmax[i] - o[i]) / d[i];
                    if (a > b) { double tmp = a; a = b; b = tmp; }
                    if (a > t0) t0 = a;
                    if (b < t1) t1 = b;
                }
                if (ok && t0 <= t1) hit_min(t0, &best);
            }
            /* camera-frame ray has z=1 => parameter == z-depth (metres) */
            double mm = best < 1e29 ? floor(best * 1000.0 + 0.5) : 0.0;
            if (mm > 65535.0) mm = 0.0;
            depth_mm[(size_t)y * w + x] = (float)mm;
        }
}

void kfo_fill_const_depth_mm(int w, int h, float mm, float *depth_mm)
{
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i) depth_mm[i] = mm;
}

/* ======================================================================
 * Whole pipeline -- kinectfusion.cpp:9-27, 48-141, 167-190
 * ====================================================================== */
#define KFO_MAX_LEVELS 8
typedef struct
{
    float *d[KFO_MAX_LEVELS], *v[KFO_MAX_LEVELS], *n[KFO_MAX_LEVELS];
} kfo_frame;

struct kfo_kinfu
{
    kfo_intr intr;
    kfo_params p;
    kfo_volume_desc vd;
    int16_t *vol;
    kfo_frame cur, prev;
    float *scratch;
    int frame_count;
    float *poses; /* 12 floats each */
    int n_poses, cap_poses;
    int64_t last_updated, last_raysteps;
    double times[4];
};

/* kinectfusion.cpp:167-190 */
void kfo_default_params(kfo_params *p, int dims)
{
    memset(p, 0, sizeof(*p));
    p->pyramid_height = 3;
    p->bfilter_color_sigma = 10;
    p->bfilter_spatial_sigma = 10;
    p->bfilter_kernel_size = 5;
    p->dfilter_dist = 5.f;
    p->icp_angle_threshold = 30.f;
    p->icp_dist_threshold = 0.015f;
    p->icp_iter_count[0] = 4; p->icp_iter_count[1] = 5; p->icp_iter_count[2] = 10;
    for (int i = 0; i < 3; ++i) { p->volu_dims[i] = dims; p->volu_range[i] = 3.f; }
    p->volu_trun_dist = 2.1f * p->volu_range[0] / (float)p->volu_dims[0];
    kfo_pose_identity(p->volu_pose);
    p->volu_pose[3] = -p->volu_range[0] / 2; p->volu_pose[7] = -p->volu_range[1] / 2; p->volu_pose[11] = 0.5f;
    p->tsdf_max_weight = 64;
    p->compat_icp_rows = 1;
    p->compat_raycast_ts_sign = 1;
}

static void frame_alloc(kfo_frame *f, const kfo_intr *k, int levels)
{
    for (int l = 0; l < levels; ++l)
    {
        kfo_intr kl;
        kfo_level_intrinsics(k, l, &kl);
        const size_t n = (size_t)kl.width * kl.height;
        f->d[l] = (float *)calloc(n, sizeof(float));
        f->v[l] = (float *)calloc(3 * n, sizeof(float));
        f->n[l] = (float *)calloc(3 * n, sizeof(float));
    }
}
static void frame_zero(kfo_frame *f, const kfo_intr *k, int levels)
{
    for (int l = 0; l < levels; ++l)
    {
        kfo_intr kl;
        kfo_level_intrinsics(k, l, &kl);
        const size_t n = (size_t)kl.width * kl.height;
        memset(f->d[l], 0, n * sizeof(float));
        memset(f->v[l], 0, 3 * n * sizeof(float));
        memset(f->n[l], 0, 3 * n * sizeof(float));
    }
}
static void frame_free(kfo_frame *f, int levels)
{
    for (int l = 0; l < levels; ++l) { free(f->d[l]); free(f->v[l]); free(f->n[l]); }
}
static void push_pose(kfo_kinfu *kf, const float p[12])
{
    if (kf->n_poses == kf->cap_poses)
    {
        kf->cap_poses = kf->cap_poses ? 2 * kf->cap_poses : 64;
        kf->poses = (float *)realloc(kf->poses, (size_t)kf->cap_poses * 12 * sizeof(float));
    }
    memcpy(kf->poses + 12 * (size_t)kf->n_poses, p, 12 * sizeof(float));
    ++kf->n_poses;
}

kfo_kinfu *kfo_kinfu_create(const kfo_intr *k, const kfo_params *p)
{
    kfo_kinfu *kf = (kfo_kinfu *)calloc(1, sizeof(kfo_kinfu));
    kf->intr = *k;
    kf->p = *p;
    for (int i = 0; i < 3; ++i)
    {
        kf->vd.dims[i] = p->volu_dims[i];
        kf->vd.range[i] = p->volu_range[i];
        kf->vd.voxel_size[i] = p->volu_range[i] / (float)p->volu_dims[i]; /* tsdf_volume.cpp:16 */
    }
    kf->vd.trunc_dist = p->volu_trun_dist;
    kf->vd.max_weight = p->tsdf_max_weight;
    const size_t nvox = (size_t)p->volu_dims[0] * p->volu_dims[1] * p->volu_dims[2];
    kf->vol = (int16_t *)calloc(2 * nvox, sizeof(int16_t));
    frame_alloc(&kf->cur, k, p->pyramid_height);
    frame_alloc(&kf->prev, k, p->pyramid_height);
    kf->scratch = (float *)calloc((size_t)k->width * k->height, sizeof(float));
    kfo_kinfu_reset(kf);
    return kf;
}
void kfo_kinfu_destroy(kfo_kinfu *kf)
{
    if (!kf) return;
    frame_free(&kf->cur, kf->p.pyramid_height);
    frame_free(&kf->prev, kf->p.pyramid_height);
    free(kf->vol); free(kf->scratch); free(kf->poses); free(kf);
}
/* kinectfusion.cpp:133-141 */
void kfo_kinfu_reset(kfo_kinfu *kf)
{
    kf->frame_count = 1;
    frame_zero(&kf->cur, &kf->intr, kf->p.pyramid_height);
    frame_zero(&kf->prev, &kf->intr, kf->p.pyramid_height);
    const size_t nvox = (size_t)kf->vd.dims[0] * kf->vd.dims[1] * kf->vd.dims[2];
    memset(kf->vol, 0, 2 * nvox * sizeof(int16_t));
    kf->n_poses = 0;
    float I[12];
    kfo_pose_identity(I);
    push_pose(kf, I);
}

/* kinectfusion.cpp:48-76 */
static void image_process(kfo_kinfu *kf, const float *depth_mm)
{
    const int L = kf->p.pyramid_height;
    kfo_intr kl[KFO_MAX_LEVELS];
    for (int l = 0; l < L; ++l) kfo_level_intrinsics(&kf->intr, l, &kl[l]);
    memcpy(kf->cur.d[0], depth_mm, (size_t)kl[0].width * kl[0].height * sizeof(float));
    for (int l = 1; l < L; ++l)
        kfo_pyrdown(kf->cur.d[l - 1], kl[l - 1].width, kl[l - 1].height, kf->cur.d[l], kl[l].width, kl[l].height);
    for (int l = 0; l < L; ++l)
    {
        const size_t n = (size_t)kl[l].width * kl[l].height;
        memcpy(kf->scratch, kf->cur.d[l], n * sizeof(float));
        kfo_bilateral(kf->scratch, kl[l].width, kl[l].height, kf->cur.d[l], kf->p.bfilter_kernel_size,
                      kf->p.bfilter_color_sigma, kf->p.bfilter_spatial_sigma);
        kfo_truncate(kf->cur.d[l], kl[l].width, kl[l].height, kf->p.dfilter_dist);
    }
    for (int l = 0; l < L; ++l)
    {
        kfo_vertex_map(kf->cur.d[l], &kl[l], kf->cur.v[l]);
        kfo_normal_map(kf->cur.v[l], kl[l].width, kl[l].height, kf->cur.n[l]);
    }
}

static void do_integrate(kfo_kinfu *kf, const float cam_pose[12])
{
    /* tsdf_volume.cpp:50: vol2cam = camera_pose.inv() * volume_pose */
    float inv[12], v2c[12];
    kfo_pose_inv(cam_pose, inv);
    kfo_pose_mul(inv, kf->p.volu_pose, v2c);
    kfo_integrate(kf->vol, &kf->vd, v2c, kf->cur.d[0], &kf->intr, 1, kf->vd.dims[2], &kf->last_updated);
}

/* icp_registration.cpp:16-45 */
static int rigid_transform(kfo_kinfu *kf, float rel[12])
{
    kfo_pose_identity(rel);
    const float sine = sinf(kf->p.icp_angle_threshold * 0.017453293f); /* icp_registration.cpp:5, types.hpp:81 */
    for (int level = kf->p.pyramid_height - 1; level >= 0; --level)
    {
        kfo_intr kl;
        kfo_level_intrinsics(&kf->intr, level, &kl);
        for (int i = 0; i < kf->p.icp_iter_count[level]; ++i)
        {
            double ab[27], x[6];
            kfo_icp_accumulate(kf->cur.v[level], kf->cur.n[level], kf->prev.v[level], kf->prev.n[level], &kl, rel,
                               kf->p.icp_dist_threshold, sine, kf->p.compat_icp_rows, ab, NULL);
            if (kfo_icp_solve(ab, x)) return 1;
            kfo_pose_apply_increment(rel, x);
        }
    }
    return 0;
}

/* kinectfusion.cpp:78-127 */
int kfo_kinfu_pipeline(kfo_kinfu *kf, const float *depth_mm)
{
    const int L = kf->p.pyramid_height;
    double t0 = now_s();
    kf->times[0] = kf->times[1] = kf->times[2] = kf->times[3] = 0.0;
    image_process(kf, depth_mm);
    kf->times[0] = now_s() - t0;
    if (kf->frame_count == 1)
    {
        t0 = now_s();
        do_integrate(kf, kf->poses + 12 * (size_t)(kf->n_poses - 1));
        kf->times[2] = now_s() - t0;
        for (int l = 0; l < L; ++l)
        {
            float *tmp = kf->cur.v[l]; kf->cur.v[l] = kf->prev.v[l]; kf->prev.v[l] = tmp;
            tmp = kf->cur.n[l]; kf->cur.n[l] = kf->prev.n[l]; kf->prev.n[l] = tmp;
        }
        kf->frame_count++;
        /* cframe->reset() (kinectfusion.cpp:91) only zeroes buffers that are fully rewritten next frame */
        return 0;
    }
    float rel[12];
    t0 = now_s();
    const int fail = rigid_transform(kf, rel);
    kf->times[1] = now_s() - t0;
    if (fail) { kfo_kinfu_reset(kf); return 1; }
    float glob[12];
    kfo_pose_mul(kf->poses + 12 * (size_t)(kf->n_poses - 1), rel, glob);
    push_pose(kf, glob);
    t0 = now_s();
    do_integrate(kf, glob);
    kf->times[2] = now_s() - t0;
    /* raycast: cam2vol = volume_pose.inv() * camera_pose (tsdf_volume.cpp:59-61) */
    t0 = now_s();
    float vinv[12], c2v[12], rinv[9];
    kfo_pose_inv(kf->p.volu_pose, vinv);
    kfo_pose_mul(vinv, glob, c2v);
    kfo_rot_inv(c2v, rinv);
    frame_zero(&kf->prev, &kf->intr, L);
    kfo_raycast(kf->vol, &kf->vd, c2v, rinv, &kf->intr, kf->prev.v[0], kf->prev.n[0], &kf->last_raysteps,
                kf->p.compat_raycast_ts_sign);
    for (int l = 1; l < L; ++l)
    {
        kfo_intr kb;
        kfo_level_intrinsics(&kf->intr, l - 1, &kb);
        kfo_resize_maps(kf->prev.v[l - 1], kf->prev.n[l - 1], kb.width, kb.height, kf->prev.v[l], kf->prev.n[l]);
    }
    kf->times[3] = now_s() - t0;
    kf->frame_count++;
    /* cframe->reset() (kinectfusion.cpp:126) is a no-op for observers: every map is rewritten next frame */
    return 0;
}

int kfo_kinfu_frame_count(const kfo_kinfu *kf) { return kf->frame_count; }
int kfo_kinfu_num_poses(const kfo_kinfu *kf) { return kf->n_poses; }
void kfo_kinfu_get_pose(const kfo_kinfu *kf, int idx, float pose12[12])
{
    if (idx < 0 || idx >= kf->n_poses) idx = kf->n_poses - 1;
    memcpy(pose12, kf->poses + 12 * (size_t)idx, 12 * sizeof(float));
}
int16_t *kfo_kinfu_volume(kfo_kinfu *kf) { return kf->vol; }
const float *kfo_kinfu_cur_depth(const kfo_kinfu *kf, int level) { return kf->cur.d[level]; }
const float *kfo_kinfu_cur_vmap(const kfo_kinfu *kf, int level) { return kf->cur.v[level]; }
const float *kfo_kinfu_cur_nmap(const kfo_kinfu *kf, int level) { return kf->cur.n[level]; }
const float *kfo_kinfu_prev_vmap(const kfo_kinfu *kf, int level) { return kf->prev.v[level]; }
const float *kfo_kinfu_prev_nmap(const kfo_kinfu *kf, int level) { return kf->prev.n[level]; }
int64_t kfo_kinfu_last_updated(const kfo_kinfu *kf) { return kf->last_updated; }
int64_t kfo_kinfu_last_raysteps(const kfo_kinfu *kf) { return kf->last_raysteps; }
void kfo_kinfu_last_times(const kfo_kinfu *kf, double t4[4]) { memcpy(t4, kf->times, sizeof(kf->times)); }
