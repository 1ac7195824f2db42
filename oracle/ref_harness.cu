// ref_harness.cu -- extern "C" entry points that drive the reference's own CUDA launchers
// (kf::device::*, compiled unmodified from /root/reference/kfusion/src/*.cu for sm_100a) on
// host-provided inputs.  TEST INFRASTRUCTURE ONLY: built by oracle/Makefile into
// oracle/_ref/libkf_ref.so and used by tests/test_ref_ab.py and bench.py --impl reference as the
// A/B checker and the reported reference baseline.  The product never links it.
//
// Layout notes: the reference voxel is 8 bytes {short tsdf; short weight; uchar3 rgb; pad}
// (device_types.hpp:51-56); hosts exchange packed int16 pairs and the harness expands/compacts.
#include <device_types.hpp>
#include <device_utils.cuh>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace kf;
using namespace kf::device;
using cv::cuda::GpuMat;

namespace
{
struct DevBuf
{
    void *p = nullptr;
    size_t bytes = 0;
    explicit DevBuf(size_t b) : bytes(b) { if (cudaMalloc(&p, b + 4096) != cudaSuccess) p = nullptr; else cudaMemset(p, 0, b + 4096); }
    ~DevBuf() { if (p) cudaFree(p); }
    void up(const void *h) { cudaMemcpy(p, h, bytes, cudaMemcpyHostToDevice); }
    void down(void *h) { cudaMemcpy(h, p, bytes, cudaMemcpyDeviceToHost); }
};

Intrs make_intrs(int w, int h, float fx, float fy, float cx, float cy)
{
    Intrinsics k;
    k.width = w; k.height = h; k.fx = fx; k.fy = fy; k.cx = cx; k.cy = cy;
    return Intrs(k);
}
PoseT make_poset(const float p[12])
{
    PoseT o;
    // PoseR stores COLUMNS (device_types.hpp:152-162): data[i] = (m(0,i), m(1,i), m(2,i))
    for (int i = 0; i < 3; ++i) { o.R.data[i].x = p[0 + i]; o.R.data[i].y = p[4 + i]; o.R.data[i].z = p[8 + i]; }
    o.t = make_float3(p[3], p[7], p[11]);
    return o;
}
PoseR make_poser9(const float m[9])
{
    PoseR o;
    for (int i = 0; i < 3; ++i) { o.data[i].x = m[0 + i]; o.data[i].y = m[3 + i]; o.data[i].z = m[6 + i]; }
    return o;
}
Volume make_volume(void *data, const int dims[3], const float range[3], float trunc)
{
    const int3 d = make_int3(dims[0], dims[1], dims[2]);
    const float3 r = make_float3(range[0], range[1], range[2]);
    const float3 vs = make_float3(range[0] / dims[0], range[1] / dims[1], range[2] / dims[2]); // tsdf_volume.cpp:16
    Volume v((Voxel *)data, d, r, vs);
    v.trun_dist = trunc;
    return v;
}
void expand(const int16_t *pairs, size_t n, std::vector<Voxel> &out)
{
    out.resize(n);
    memset(out.data(), 0, n * sizeof(Voxel));
    for (size_t i = 0; i < n; ++i) { out[i].tsdf = pairs[2 * i]; out[i].weight = pairs[2 * i + 1]; }
}
cudaEvent_t ev0, ev1;
bool ev_init = false;
void tic() { if (!ev_init) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); ev_init = true; } cudaEventRecord(ev0, 0); }
float toc() { cudaEventRecord(ev1, 0); cudaEventSynchronize(ev1); float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1); return ms; }
} // namespace

extern "C" {

int ref_sizeof_voxel(void) { return (int)sizeof(Voxel); }

// ---- persistent reference volume (8-byte voxels) for integrate / raycast / extract ------------------
struct RefVolume
{
    DevBuf *buf;
    int dims[3];
    float range[3];
    float trunc;
};
void *ref_volume_create(const int dims[3], const float range[3], float trunc)
{
    RefVolume *v = new RefVolume();
    memcpy(v->dims, dims, sizeof(v->dims));
    memcpy(v->range, range, sizeof(v->range));
    v->trunc = trunc;
    v->buf = new DevBuf((size_t)dims[0] * dims[1] * dims[2] * sizeof(Voxel));
    if (!v->buf->p) { delete v->buf; delete v; return nullptr; }
    return v;
}
void ref_volume_destroy(void *h) { RefVolume *v = (RefVolume *)h; delete v->buf; delete v; }
void ref_volume_upload(void *h, const int16_t *pairs)
{
    RefVolume *v = (RefVolume *)h;
    std::vector<Voxel> tmp;
    expand(pairs, (size_t)v->dims[0] * v->dims[1] * v->dims[2], tmp);
    v->buf->up(tmp.data());
}
void ref_volume_download(void *h, int16_t *pairs)
{
    RefVolume *v = (RefVolume *)h;
    const size_t n = (size_t)v->dims[0] * v->dims[1] * v->dims[2];
    std::vector<Voxel> tmp(n);
    v->buf->down(tmp.data());
    for (size_t i = 0; i < n; ++i) { pairs[2 * i] = tmp[i].tsdf; pairs[2 * i + 1] = tmp[i].weight; }
}
// device::integrate (tsdf_volume.cu:103-111); returns device ms of the launch (+ its own sync)
float ref_integrate(void *h, const float vol2cam12[12], const float *depth_m, int w, int hgt, float fx, float fy, float cx, float cy, int reps)
{
    RefVolume *v = (RefVolume *)h;
    DevBuf d((size_t)w * hgt * 4), c((size_t)w * hgt * 3);
    d.up(depth_m);
    GpuMat dm(hgt, w, 4, d.p), cm(hgt, w, 3, c.p);
    Volume vol = make_volume(v->buf->p, v->dims, v->range, v->trunc);
    float ms = 0;
    for (int r = 0; r < reps; ++r)
    {
        tic();
        integrate(make_intrs(w, hgt, fx, fy, cx, cy), make_poset(vol2cam12), vol, dm, cm);
        ms = toc();
    }
    return ms;
}
// device::raycast (tsdf_volume.cu:264-273) into zeroed maps (pframe->reset(), kinectfusion.cpp:112)
float ref_raycast(void *h, const float cam2vol12[12], const float rinv9[9], int w, int hgt, float fx, float fy, float cx, float cy,
                  float *vmap3, float *nmap3, int reps)
{
    RefVolume *v = (RefVolume *)h;
    DevBuf vm((size_t)w * hgt * 12), nm((size_t)w * hgt * 12);
    GpuMat gv(hgt, w, 12, vm.p), gn(hgt, w, 12, nm.p);
    Volume vol = make_volume(v->buf->p, v->dims, v->range, v->trunc);
    float ms = 0;
    for (int r = 0; r < reps; ++r)
    {
        gv.setTo(0); gn.setTo(0);
        tic();
        raycast(make_intrs(w, hgt, fx, fy, cx, cy), make_poset(cam2vol12), make_poser9(rinv9), vol, gv, gn);
        ms = toc();
    }
    vm.down(vmap3); nm.down(nmap3);
    return ms;
}
// device::extract_points (tsdf_volume.cu:483-499)
long ref_extract_points(void *h, const float volpose12[12], float *points3, long cap)
{
    RefVolume *v = (RefVolume *)h;
    DevBuf pts((size_t)cap * 12);
    Volume vol = make_volume(v->buf->p, v->dims, v->range, v->trunc);
    cv::cuda::PtrSz<Point3> arr;
    arr.data = (Point3 *)pts.p;
    arr.size = (size_t)cap;
    const size_t n = extract_points(vol, arr, make_poset(volpose12));
    if (n) cudaMemcpy(points3, pts.p, n * 12, cudaMemcpyDeviceToHost);
    return (long)n;
}

// ---- image kernels (image_process.cu) ---------------------------------------------------------------
void ref_depth_truncation(float *depth_inout, int w, int h, float max_dist)
{
    DevBuf d((size_t)w * h * 4);
    d.up(depth_inout);
    GpuMat dm(h, w, 4, d.p);
    depthTruncation(dm, max_dist);
    cudaDeviceSynchronize();
    d.down(depth_inout);
}
void ref_vertex_normal(const float *depth_m, int w, int h, float fx, float fy, float cx, float cy, float *vmap3, float *nmap3)
{
    DevBuf d((size_t)w * h * 4), vm((size_t)w * h * 12), nm((size_t)w * h * 12);
    d.up(depth_m);
    GpuMat dm(h, w, 4, d.p), gv(h, w, 12, vm.p), gn(h, w, 12, nm.p);
    gv.setTo(0); gn.setTo(0);
    getVertexmap(dm, gv, make_intrs(w, h, fx, fy, cx, cy));
    getNormalmap(gv, gn);
    cudaDeviceSynchronize();
    vm.down(vmap3); nm.down(nmap3);
}
void ref_resize_maps(const float *vbig, const float *nbig, int bw, int bh, float *vsmall, float *nsmall)
{
    const int sw = bw >> 1, sh = bh >> 1;
    DevBuf vb((size_t)bw * bh * 12), nb((size_t)bw * bh * 12), vs((size_t)sw * sh * 12), ns((size_t)sw * sh * 12);
    vb.up(vbig); nb.up(nbig);
    GpuMat gvb(bh, bw, 12, vb.p), gnb(bh, bw, 12, nb.p), gvs(sh, sw, 12, vs.p), gns(sh, sw, 12, ns.p);
    resizePointsNormals(gvb, gnb, gvs, gns);
    cudaDeviceSynchronize();
    vs.down(vsmall); ns.down(nsmall);
}
void ref_render(const float *vmap3, const float *nmap3, int w, int h, const float eye[3], int phong, unsigned char *bgr)
{
    DevBuf vm((size_t)w * h * 12), nm((size_t)w * h * 12), cm((size_t)w * h * 3);
    vm.up(vmap3); nm.up(nmap3);
    GpuMat gv(h, w, 12, vm.p), gn(h, w, 12, nm.p), gc(h, w, 3, cm.p);
    gc.setTo(0);
    if (phong) renderPhong(make_float3(eye[0], eye[1], eye[2]), gv, gn, gc);
    else renderNormals(gn, gc);
    cudaDeviceSynchronize();
    cm.down(bgr);
}

// ---- ICP (rigid_icp.cu:135-169) -----------------------------------------------------------------------
// A: 36 doubles row-major, b: 6 doubles.  Returns device+host ms of the last call.
float ref_rigid_icp(const float *cur_v3, const float *cur_n3, const float *pre_v3, const float *pre_n3, int w, int h,
                    float fx, float fy, float cx, float cy, const float pose12[12], float dist_thres, float sine_thres,
                    double *A36, double *b6, int reps)
{
    const size_t nb = (size_t)w * h * 12;
    DevBuf cv_(nb), cn(nb), pv(nb), pn(nb);
    cv_.up(cur_v3); cn.up(cur_n3); pv.up(pre_v3); pn.up(pre_n3);
    ICP helper(dist_thres, sine_thres);
    helper.cur_vmap = cv::cuda::PtrStep<float3>((float3 *)cv_.p, (size_t)w * 12);
    helper.cur_nmap = cv::cuda::PtrStep<float3>((float3 *)cn.p, (size_t)w * 12);
    helper.pre_vmap = cv::cuda::PtrStep<float3>((float3 *)pv.p, (size_t)w * 12);
    helper.pre_nmap = cv::cuda::PtrStep<float3>((float3 *)pn.p, (size_t)w * 12);
    helper.setIntrs(make_intrs(w, h, fx, fy, cx, cy), w, h);
    helper.curpose = make_poset(pose12);
    cv::Matx66d A;
    cv::Vec6d b;
    float ms = 0;
    for (int r = 0; r < reps; ++r)
    {
        tic();
        rigidICP(helper, A, b);
        ms = toc();
    }
    for (int i = 0; i < 36; ++i) A36[i] = A.val[i];
    for (int i = 0; i < 6; ++i) b6[i] = b.val[i];
    return ms;
}

int ref_last_cuda_error(void) { return (int)cudaGetLastError(); }

} // extern "C"

// ---- whole-frame driver over the reference's own kernels (bench.py --impl reference) ------------------------
// Mirrors kf::kinectfusion::pipeline (kfusion/src/kinectfusion.cpp:48-127) and ICPRegistration::rigidTransform
// (icp_registration.cpp:16-45) with the reference's launchers for every stage it owns (depthTruncation,
// getVertexmap, getNormalmap, rigidICP incl. its per-iteration cudaMalloc/cudaFree/memcpy, integrate incl. its
// cudaDeviceSynchronize, raycast, resizePointsNormals) on its 8-byte voxel.  The two OpenCV-CUDA calls the
// reference makes (cv::cuda::pyrDown, cv::cuda::bilateralFilter) are not vendored; they are stood in for by the
// two plain kernels below, written here from the published algorithm -- baseline plumbing, not
// product code.  Host algebra (6x6 solve, Rodrigues, 4x4 products) in double/float like OpenCV core.
namespace
{
__device__ __forceinline__ int refl101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}
__global__ void base_pyrdown(const float *src, int w, int h, float *dst, int dw, int dh)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh) return;
    const float wt[5] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
    float acc = 0.f;
    for (int k = 0; k < 5; ++k)
    {
        const int c = refl101(2 * x - 2 + k, w);
        float s = 0.f;
        for (int j = 0; j < 5; ++j) s += wt[j] * src[(size_t)refl101(2 * y - 2 + j, h) * w + c];
        acc += wt[k] * s;
    }
    dst[(size_t)y * dw + x] = acc;
}
__global__ void base_bilateral(const float *src, float *dst, int w, int h, int r, float ss, float sc)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float center = src[(size_t)y * w + x];
    float s1 = 0.f, s2 = 0.f;
    for (int cy = y - r; cy <= y + r; ++cy)
        for (int cx = x - r; cx <= x + r; ++cx)
        {
            const float sp = (float)((cx - x) * (cx - x) + (cy - y) * (cy - y));
            if (sp > (float)(r * r)) continue;
            const float v = src[(size_t)refl101(cy, h) * w + refl101(cx, w)];
            const float d = fabsf(v - center);
            const float wgt = expf(sp * ss + d * d * sc);
            s1 += wgt * v; s2 += wgt;
        }
    dst[(size_t)y * w + x] = s1 / s2;
}

struct Mat4 { float m[16]; };
Mat4 ident4() { Mat4 a; memset(a.m, 0, sizeof(a.m)); a.m[0] = a.m[5] = a.m[10] = a.m[15] = 1.f; return a; }
Mat4 mul4(const Mat4 &a, const Mat4 &b)
{
    Mat4 c;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
        {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a.m[4 * i + k] * b.m[4 * k + j];
            c.m[4 * i + j] = s;
        }
    return c;
}
Mat4 inv_rigid_general(const Mat4 &a) // full 3x3 inverse of the rotation block (Affine3f::inv is a 4x4 inverse)
{
    const float *m = a.m;
    const double a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[4], a11 = m[5], a12 = m[6], a20 = m[8], a21 = m[9], a22 = m[10];
    const double det = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
    const double id = 1.0 / det;
    double r[9] = {(a11 * a22 - a12 * a21) * id, (a02 * a21 - a01 * a22) * id, (a01 * a12 - a02 * a11) * id,
                   (a12 * a20 - a10 * a22) * id, (a00 * a22 - a02 * a20) * id, (a02 * a10 - a00 * a12) * id,
                   (a10 * a21 - a11 * a20) * id, (a01 * a20 - a00 * a21) * id, (a00 * a11 - a01 * a10) * id};
    Mat4 o = ident4();
    for (int i = 0; i < 3; ++i)
    {
        for (int j = 0; j < 3; ++j) o.m[4 * i + j] = (float)r[3 * i + j];
        o.m[4 * i + 3] = (float)(-(r[3 * i] * m[3] + r[3 * i + 1] * m[7] + r[3 * i + 2] * m[11]));
    }
    return o;
}
Mat4 from_rvec_t(const double x[6])
{
    Mat4 o = ident4();
    const double rx = (float)x[0], ry = (float)x[1], rz = (float)x[2];
    const double th = sqrt(rx * rx + ry * ry + rz * rz);
    if (th >= 2.220446049250313e-16)
    {
        const double c = cos(th), s = sin(th), c1 = 1.0 - c, it = 1.0 / th;
        const double ux = rx * it, uy = ry * it, uz = rz * it;
        const double R[9] = {c + c1 * ux * ux, c1 * ux * uy - s * uz, c1 * ux * uz + s * uy,
                             c1 * ux * uy + s * uz, c + c1 * uy * uy, c1 * uy * uz - s * ux,
                             c1 * ux * uz - s * uy, c1 * uy * uz + s * ux, c + c1 * uz * uz};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.m[4 * i + j] = (float)R[3 * i + j];
    }
    o.m[3] = (float)x[3]; o.m[7] = (float)x[4]; o.m[11] = (float)x[5];
    return o;
}
bool solve6(const cv::Matx66d &A, const cv::Vec6d &b, double x[6]) // det guard + LU solve (icp_registration.cpp:35-39)
{
    double M[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = A.val[6 * i + j]; M[i][6] = b.val[i]; }
    double det = 1.0;
    for (int c = 0; c < 6; ++c)
    {
        int p = c;
        for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
        if (M[p][c] == 0.0 || M[p][c] != M[p][c]) return false;
        if (p != c) { for (int j = 0; j < 7; ++j) { double t = M[c][j]; M[c][j] = M[p][j]; M[p][j] = t; } det = -det; }
        det *= M[c][c];
        for (int r = c + 1; r < 6; ++r)
        {
            const double f = M[r][c] / M[c][c];
            for (int j = c; j < 7; ++j) M[r][j] -= f * M[c][j];
        }
    }
    if (fabs(det) < 1e-15 || det != det) return false;
    for (int i = 5; i >= 0; --i) { double s = M[i][6]; for (int q = i + 1; q < 6; ++q) s -= M[i][q] * x[q]; x[i] = s / M[i][i]; }
    return true;
}
void to12(const Mat4 &a, float p[12]) { memcpy(p, a.m, 12 * sizeof(float)); }

struct RefFrame { DevBuf *d[8], *v[8], *n[8]; };
} // namespace

struct RefKinfu
{
    int w, h, levels, dims[3];
    float fx, fy, cx, cy, range[3], trunc;
    int iters[8];
    DevBuf *raw[8], *tmp[8];
    RefFrame cur, prev;
    DevBuf *color, *vol;
    Mat4 volpose, pose;
    int frame_count;
};

extern "C" {

void *ref_kinfu_create(int w, int h, float fx, float fy, float cx, float cy, int dim, float range, const float volpose12[12])
{
    RefKinfu *k = new RefKinfu();
    k->w = w; k->h = h; k->fx = fx; k->fy = fy; k->cx = cx; k->cy = cy; k->levels = 3;
    for (int i = 0; i < 3; ++i) { k->dims[i] = dim; k->range[i] = range; }
    k->trunc = 2.1f * range / (float)dim;                 // kinectfusion.cpp:183
    k->iters[0] = 4; k->iters[1] = 5; k->iters[2] = 10;   // kinectfusion.cpp:176
    for (int l = 0; l < 3; ++l)
    {
        const size_t n = (size_t)(w >> l) * (h >> l);
        k->raw[l] = new DevBuf(n * 4); k->tmp[l] = new DevBuf(n * 4);
        k->cur.d[l] = new DevBuf(n * 4); k->cur.v[l] = new DevBuf(n * 12); k->cur.n[l] = new DevBuf(n * 12);
        k->prev.d[l] = new DevBuf(n * 4); k->prev.v[l] = new DevBuf(n * 12); k->prev.n[l] = new DevBuf(n * 12);
    }
    k->color = new DevBuf((size_t)w * h * 3);
    k->vol = new DevBuf((size_t)dim * dim * dim * sizeof(Voxel));
    if (!k->vol->p) return nullptr;
    k->volpose = ident4();
    memcpy(k->volpose.m, volpose12, 12 * sizeof(float));
    k->pose = ident4();
    k->frame_count = 1;
    return k;
}
void ref_kinfu_reset(void *hnd)
{
    RefKinfu *k = (RefKinfu *)hnd;
    cudaMemset(k->vol->p, 0, k->vol->bytes); // the intent of device::resetVolume (its fixed grid is a reference bug)
    k->pose = ident4();
    k->frame_count = 1;
}
void ref_kinfu_destroy(void *hnd)
{
    RefKinfu *k = (RefKinfu *)hnd;
    for (int l = 0; l < 3; ++l)
    {
        delete k->raw[l]; delete k->tmp[l];
        delete k->cur.d[l]; delete k->cur.v[l]; delete k->cur.n[l];
        delete k->prev.d[l]; delete k->prev.v[l]; delete k->prev.n[l];
    }
    delete k->color; delete k->vol; delete k;
}
void ref_kinfu_get_pose(void *hnd, float pose12[12]) { to12(((RefKinfu *)hnd)->pose, pose12); }

// one frame; `depth_mm` may be a host (pinned or pageable) or device pointer.  Returns 0 ok, 1 tracking failure.
int ref_kinfu_pipeline(void *hnd, const float *depth_mm)
{
    RefKinfu *k = (RefKinfu *)hnd;
    const int w = k->w, h = k->h;
    // ---- imageProcess (kinectfusion.cpp:48-76)
    cudaMemcpy(k->raw[0]->p, depth_mm, (size_t)w * h * 4, cudaMemcpyDefault);                 // GpuMat::upload (:50)
    for (int l = 1; l < 3; ++l)
    {
        const int sw = w >> (l - 1), sh = h >> (l - 1), dw = w >> l, dh = h >> l;
        dim3 b(32, 8), g((dw + 31) / 32, (dh + 7) / 8);
        base_pyrdown<<<g, b>>>((const float *)k->raw[l - 1]->p, sw, sh, (float *)k->raw[l]->p, dw, dh); // cv::cuda::pyrDown (:55)
    }
    for (int l = 0; l < 3; ++l)
    {
        const int lw = w >> l, lh = h >> l;
        dim3 b(32, 8), g((lw + 31) / 32, (lh + 7) / 8);
        base_bilateral<<<g, b>>>((const float *)k->raw[l]->p, (float *)k->cur.d[l]->p, lw, lh, 2, -0.005f, -0.005f); // (:60-64)
        GpuMat dm(lh, lw, 4, k->cur.d[l]->p), gv(lh, lw, 12, k->cur.v[l]->p), gn(lh, lw, 12, k->cur.n[l]->p);
        depthTruncation(dm, 5.f);                                                            // (:66)
        Intrinsics ki;
        const float sc = powf(0.5f, (float)l);
        ki.width = lw; ki.height = lh; ki.fx = k->fx * sc; ki.fy = k->fy * sc;
        ki.cx = l ? (k->cx + 0.5f) * sc - 0.5f : k->cx; ki.cy = l ? (k->cy + 0.5f) * sc - 0.5f : k->cy;
        gv.setTo(0); gn.setTo(0);                                                            // Frame::reset (:126, types.hpp:53-62)
        getVertexmap(dm, gv, Intrs(ki));                                                     // (:73)
        getNormalmap(gv, gn);                                                                // (:74)
    }
    const Intrs K0 = make_intrs(w, h, k->fx, k->fy, k->cx, k->cy);
    Volume vol = make_volume(k->vol->p, k->dims, k->range, k->trunc);
    GpuMat d0(h, w, 4, k->cur.d[0]->p), c0(h, w, 3, k->color->p);
    float p12[12];
    if (k->frame_count == 1)
    {
        to12(mul4(inv_rigid_general(k->pose), k->volpose), p12);
        integrate(K0, make_poset(p12), vol, d0, c0);                                         // (:86)
        for (int l = 0; l < 3; ++l) { std::swap(k->cur.v[l], k->prev.v[l]); std::swap(k->cur.n[l], k->prev.n[l]); } // (:88-89)
        k->frame_count++;
        return 0;
    }
    // ---- ICPRegistration::rigidTransform (icp_registration.cpp:16-45)
    Mat4 rel = ident4();
    const float sine_thres = sinf(30.f * 0.017453293f);
    for (int level = 2; level >= 0; --level)
    {
        const int lw = w >> level, lh = h >> level;
        Intrinsics ki;
        const float sc = powf(0.5f, (float)level);
        ki.width = lw; ki.height = lh; ki.fx = k->fx * sc; ki.fy = k->fy * sc;
        ki.cx = level ? (k->cx + 0.5f) * sc - 0.5f : k->cx; ki.cy = level ? (k->cy + 0.5f) * sc - 0.5f : k->cy;
        ICP helper(0.015f, sine_thres);
        helper.cur_vmap = cv::cuda::PtrStep<float3>((float3 *)k->cur.v[level]->p, (size_t)lw * 12);
        helper.cur_nmap = cv::cuda::PtrStep<float3>((float3 *)k->cur.n[level]->p, (size_t)lw * 12);
        helper.pre_vmap = cv::cuda::PtrStep<float3>((float3 *)k->prev.v[level]->p, (size_t)lw * 12);
        helper.pre_nmap = cv::cuda::PtrStep<float3>((float3 *)k->prev.n[level]->p, (size_t)lw * 12);
        helper.setIntrs(Intrs(ki), lw, lh);
        for (int it = 0; it < k->iters[level]; ++it)
        {
            to12(rel, p12);
            helper.curpose = make_poset(p12);
            cv::Matx66d A;
            cv::Vec6d b;
            rigidICP(helper, A, b);                                                          // (:33)
            double x[6];
            if (!solve6(A, b, x)) { ref_kinfu_reset(k); return 1; }
            rel = mul4(rel, from_rvec_t(x));                                                 // (:41-42)
        }
    }
    k->pose = mul4(k->pose, rel);                                                            // kinectfusion.cpp:104
    to12(mul4(inv_rigid_general(k->pose), k->volpose), p12);
    integrate(K0, make_poset(p12), vol, d0, c0);                                             // (:107)
    const Mat4 c2v = mul4(inv_rigid_general(k->volpose), k->pose);
    Mat4 rot = c2v; rot.m[3] = rot.m[7] = rot.m[11] = 0.f;
    const Mat4 rinv = inv_rigid_general(rot);
    float r9[9] = {rinv.m[0], rinv.m[1], rinv.m[2], rinv.m[4], rinv.m[5], rinv.m[6], rinv.m[8], rinv.m[9], rinv.m[10]};
    to12(c2v, p12);
    GpuMat pv(h, w, 12, k->prev.v[0]->p), pn(h, w, 12, k->prev.n[0]->p);
    pv.setTo(0); pn.setTo(0);                                                                // pframe->reset() (:112)
    raycast(K0, make_poset(p12), make_poser9(r9), vol, pv, pn);                              // (:113)
    for (int l = 1; l < 3; ++l)
    {
        const int bw = w >> (l - 1), bh = h >> (l - 1), sw = w >> l, sh = h >> l;
        GpuMat gvb(bh, bw, 12, k->prev.v[l - 1]->p), gnb(bh, bw, 12, k->prev.n[l - 1]->p);
        GpuMat gvs(sh, sw, 12, k->prev.v[l]->p), gns(sh, sw, 12, k->prev.n[l]->p);
        gvs.setTo(0); gns.setTo(0);
        resizePointsNormals(gvb, gnb, gvs, gns);                                             // (:116)
    }
    k->frame_count++;
    return 0;
}
void ref_device_sync(void) { cudaDeviceSynchronize(); }
float ref_event_ms(int which) // 0: tic now, 1: toc -> ms since tic (legacy default stream)
{
    if (which == 0) { tic(); return 0.f; }
    return toc();
}

} // extern "C"

// ---- lock-step checking (tests/test_ref_full.py): put the reference loop into the state of another pipeline ----
// vol_packed: device pointer to the product's volume (kfb_device_ptr(ctx, 0): packed voxels, brick-major);
// vmap4 / nmap4: device pointers to level-0 model maps as float4 (w unused); pose12: camera pose.  Everything is
// converted on the device into the reference's own layouts (8-byte Voxel, float3 maps); the coarser model levels are
// rebuilt with the reference's resizePointsNormals, as its frame loop does after every raycast.
namespace
{
// src: the product's brick-major volume (8x8x8 voxel bricks of packed {int16 tsdf, int16 weight}, bricks and voxels
// in x, y, z order, dims padded to whole bricks); dst: the reference's linear array of 8-byte voxels
__global__ void unpack_volume(const uint32_t *src, Voxel *dst, int X, int Y, int Z)
{
    const size_t n = (size_t)X * Y * Z;
    const int bx = (X + 7) >> 3, by = (Y + 7) >> 3;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        const int x = (int)(i % X), y = (int)((i / X) % Y), z = (int)(i / ((size_t)X * Y));
        const size_t b = ((size_t)(((z >> 3) * by + (y >> 3)) * bx + (x >> 3)) << 9) + (size_t)(((z & 7) << 6) | ((y & 7) << 3) | (x & 7));
        const uint32_t w = src[b];
        Voxel v;
        memset(&v, 0, sizeof(v));
        v.tsdf = (short)(w & 0xffffu);
        v.weight = (short)(w >> 16);
        dst[i] = v;
    }
}
__global__ void f4_to_f3(const float4 *src, float3 *dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float4 v = src[i]; dst[i] = make_float3(v.x, v.y, v.z); }
}
} // namespace
extern "C" void ref_kinfu_sync_from(void *hnd, const void *vol_packed, const void *vmap4, const void *nmap4, const float pose12[12], int frame_count)
{
    RefKinfu *k = (RefKinfu *)hnd;
    unpack_volume<<<148 * 8, 256>>>((const uint32_t *)vol_packed, (Voxel *)k->vol->p, k->dims[0], k->dims[1], k->dims[2]);
    const int np = k->w * k->h;
    f4_to_f3<<<(np + 255) / 256, 256>>>((const float4 *)vmap4, (float3 *)k->prev.v[0]->p, np);
    f4_to_f3<<<(np + 255) / 256, 256>>>((const float4 *)nmap4, (float3 *)k->prev.n[0]->p, np);
    for (int l = 1; l < 3; ++l)
    {
        const int bw = k->w >> (l - 1), bh = k->h >> (l - 1), sw = k->w >> l, sh = k->h >> l;
        GpuMat gvb(bh, bw, 12, k->prev.v[l - 1]->p), gnb(bh, bw, 12, k->prev.n[l - 1]->p);
        GpuMat gvs(sh, sw, 12, k->prev.v[l]->p), gns(sh, sw, 12, k->prev.n[l]->p);
        gvs.setTo(0); gns.setTo(0);
        resizePointsNormals(gvb, gnb, gvs, gns);
    }
    k->pose = ident4();
    memcpy(k->pose.m, pose12, 12 * sizeof(float));
    k->frame_count = frame_count;
    cudaDeviceSynchronize();
}
