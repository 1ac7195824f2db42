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
