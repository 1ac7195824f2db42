// c_api.cpp -- extern "C" handle API over kf::kinectfusion so the facade can be driven from
// ctypes (tests, bench.py) exactly as a C++ application drives the reference's class.
#include <kinectfusion.h>
#include <pose_mailbox.hpp>
#include <depth_sensor.h>
#include <cstring>
#include <string>

struct kfh_params
{
    int pyramid_height;
    float dfilter_dist;
    int bfilter_kernel_size;
    float bfilter_spatial_sigma, bfilter_color_sigma;
    float icp_dist_threshold, icp_angle_threshold;
    int icp_iter_count[KFB_MAX_LEVELS];
    int volu_dims[3];
    float volu_range[3];
    float volu_pose[12];
    float volu_trun_dist;
    int tsdf_max_weight;
    int compat_icp_rows, compat_raycast_ts_sign;
    int device;
    int slab_z_begin, slab_z_end, shard_rank, shard_world;
};

static thread_local std::string g_err;

extern "C" {

void kfh_default_params(kfh_params *o, int dims)
{
    kf::kinectfuison_params p = kf::kinectfuison_params::default_params();
    std::memset(o, 0, sizeof(*o));
    o->pyramid_height = p.pyramid_height;
    o->dfilter_dist = p.dfilter_dist;
    o->bfilter_kernel_size = p.bfilter_kernel_size;
    o->bfilter_spatial_sigma = p.bfilter_spatial_sigma;
    o->bfilter_color_sigma = p.bfilter_color_sigma;
    o->icp_dist_threshold = p.icp_dist_threshold;
    o->icp_angle_threshold = p.icp_angle__threshold;
    for (size_t i = 0; i < p.icp_iter_count.size(); ++i) o->icp_iter_count[i] = p.icp_iter_count[i];
    for (int i = 0; i < 3; ++i) { o->volu_dims[i] = dims; o->volu_range[i] = p.volu_range(i); }
    p.volu_pose.to12(o->volu_pose);
    o->volu_trun_dist = 2.1f * p.volu_range(0) / (float)dims;
    o->tsdf_max_weight = p.tsdf_max_weight;
    o->compat_icp_rows = 1;
    o->compat_raycast_ts_sign = 1;
    o->device = 0;
    o->slab_z_begin = o->slab_z_end = 0;
    o->shard_rank = 0; o->shard_world = 1;
}

void *kfh_create(const kfb_intrinsics *k, const kfh_params *q)
{
    try
    {
        kf::Intrinsics intr{k->width, k->height, k->fx, k->fy, k->cx, k->cy};
        kf::kinectfuison_params p = kf::kinectfuison_params::default_params();
        p.pyramid_height = q->pyramid_height;
        p.dfilter_dist = q->dfilter_dist;
        p.bfilter_kernel_size = q->bfilter_kernel_size;
        p.bfilter_spatial_sigma = q->bfilter_spatial_sigma;
        p.bfilter_color_sigma = q->bfilter_color_sigma;
        p.icp_dist_threshold = q->icp_dist_threshold;
        p.icp_angle__threshold = q->icp_angle_threshold;
        p.icp_iter_count.assign(q->icp_iter_count, q->icp_iter_count + q->pyramid_height);
        for (int i = 0; i < 3; ++i) { p.volu_dims(i) = q->volu_dims[i]; p.volu_range(i) = q->volu_range[i]; }
        p.volu_pose = cv::Affine3f::from12(q->volu_pose);
        p.volu_trun_dist = q->volu_trun_dist;
        p.tsdf_max_weight = q->tsdf_max_weight;
        p.compat_icp_rows = q->compat_icp_rows;
        p.compat_raycast_ts_sign = q->compat_raycast_ts_sign;
        p.device = q->device;
        p.slab_z_begin = q->slab_z_begin; p.slab_z_end = q->slab_z_end;
        p.shard_rank = q->shard_rank; p.shard_world = q->shard_world > 0 ? q->shard_world : 1;
        return new kf::kinectfusion(intr, p);
    }
    catch (const std::exception &e)
    {
        g_err = e.what();
        return nullptr;
    }
}
const char *kfh_last_error(void) { return g_err.c_str(); }
void kfh_destroy(void *h) { delete static_cast<kf::kinectfusion *>(h); }
void kfh_reset(void *h) { static_cast<kf::kinectfusion *>(h)->reset(); }
/* returns 0 ok, 1 tracking failure (reset done, kinectfusion.cpp:97-102) */
int kfh_pipeline(void *h, const float *depth_mm, int width, int height)
{
    kf::kinectfusion *k = static_cast<kf::kinectfusion *>(h);
    try { k->pipeline(depth_mm, width, height); }
    catch (const std::exception &e) { g_err = e.what(); return 2; }
    return k->last_tracking_ok ? 0 : 1;
}
/* collectives of a z-slab sharded instance (kf::ShardComm) */
void kfh_set_shard_comm(void *h, int (*bcast)(float *, void *), int (*composite)(void *), void *user)
{
    kf::ShardComm c;
    c.broadcast_pose = bcast; c.composite = composite; c.user = user;
    static_cast<kf::kinectfusion *>(h)->setShardComm(c);
}
/* native pose mailbox (pose_mailbox.hpp): handle or NULL; kfh_set_pose_mailbox installs it as the instance's
 * broadcast_pose (the composite callback stays whatever kfh_set_shard_comm installed, or none with peer memory) */
void *kfh_mailbox_open(const char *name, int rank, int world)
{
    kf::PoseMailbox *m = new kf::PoseMailbox();
    if (!m->open(name, rank, world)) { g_err = m->lastError(); delete m; return nullptr; }
    return m;
}
void kfh_mailbox_close(void *m) { delete static_cast<kf::PoseMailbox *>(m); }
int kfh_mailbox_exchange(void *m, float *msg13) { return static_cast<kf::PoseMailbox *>(m)->exchange(msg13); }
void kfh_set_pose_mailbox(void *h, void *m, int (*composite)(void *), void *user)
{
    kf::ShardComm c;
    c.broadcast_pose = &kf::PoseMailbox::callback; c.pose_user = m; c.composite = composite; c.user = user;
    static_cast<kf::kinectfusion *>(h)->setShardComm(c);
}
double kfh_last_icp_us(void *h) { return static_cast<kf::kinectfusion *>(h)->last_icp_us; }
int kfh_frame_count(void *h) { return static_cast<kf::kinectfusion *>(h)->frame_count; }
int kfh_num_poses(void *h) { return (int)static_cast<kf::kinectfusion *>(h)->pose_record.size(); }
void kfh_get_pose(void *h, int idx, float pose12[12])
{
    kf::kinectfusion *k = static_cast<kf::kinectfusion *>(h);
    if (idx < 0 || idx >= (int)k->pose_record.size()) idx = (int)k->pose_record.size() - 1;
    k->pose_record[idx].to12(pose12);
}
void *kfh_context(void *h) { return static_cast<kf::kinectfusion *>(h)->context(); }
int kfh_render(void *h, int normal, uint8_t *bgr)
{
    kf::kinectfusion *k = static_cast<kf::kinectfusion *>(h);
    cv::Mat m = k->getRenderMap(normal ? kf::kinectfusion::NORMAL : kf::kinectfusion::PHONG);
    std::memcpy(bgr, m.ptr<uint8_t>(), (size_t)m.rows * m.cols * 3);
    return 0;
}
/* extracePointcloud(): returns N, copies up to cap points */
long kfh_extract_pointcloud(void *h, float *points3, long cap)
{
    kf::kinectfusion *k = static_cast<kf::kinectfusion *>(h);
    cv::Mat m = k->extracePointcloud();
    const long n = m.cols < cap ? m.cols : cap;
    if (n > 0) std::memcpy(points3, m.ptr<float>(), (size_t)n * 12);
    return m.cols;
}
int kfh_save_pointcloud(void *h, const char *path)
{
    static_cast<kf::kinectfusion *>(h)->savePointcloud(path);
    return 0;
}
/* diagnostic: wall-clock microseconds of every kfb_icp_step of one schedule (identity poses, no solve) */
int kfh_icp_probe(void *kfb_ctx_handle, const int *iters_per_level, double *us_out, int cap)
{
    kfb_ctx *ctx = static_cast<kfb_ctx *>(kfb_ctx_handle);
    const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    int sched[KFB_MAX_LEVELS] = {0}, total = 0;
    for (int l = 0; l < KFB_MAX_LEVELS; ++l) { sched[l] = iters_per_level[l]; total += sched[l]; }
    if (kfb_icp_begin(ctx, sched) != KFB_OK) return -1;
    double sums[27];
    int n = 0;
    for (int k = 0; k < total; ++k)
    {
        const auto t0 = std::chrono::steady_clock::now();
        if (kfb_icp_step(ctx, I, sums) != KFB_OK) { kfb_icp_end(ctx); return -2; }
        const std::chrono::duration<double, std::micro> dt = std::chrono::steady_clock::now() - t0;
        if (n < cap) us_out[n++] = dt.count();
    }
    kfb_icp_end(ctx);
    return n;
}
int kfh_save_volume(void *h, const char *path) { return static_cast<kf::kinectfusion *>(h)->volume()->save(path) ? 0 : 1; }
int kfh_load_volume(void *h, const char *path) { return static_cast<kf::kinectfusion *>(h)->volume()->load(path) ? 0 : 1; }
int kfh_save_poses(void *h, const char *path)
{
    kf::kinectfusion *k = static_cast<kf::kinectfusion *>(h);
    return kf::file::exportPoses(path, k->pose_record) ? 0 : 1;
}
/* out5 = fx, cx, fy, cy, depth scale */
int kfh_read_intrinsics(const char *path, float out5[5])
{
    kf::Intrinsics k{0, 0, 0.f, 0.f, 0.f, 0.f};
    if (!kf::file::readIntrinsics(path, k)) return 1;
    out5[0] = k.fx; out5[1] = k.cx; out5[2] = k.fy; out5[3] = k.cy; out5[4] = k.c;
    return 0;
}
/* dataset frame source (depth_sensor.h): open -> handle or NULL; info8 = width, height, fx, cx, fy, cy, scale, frames left */
void *kfh_sensor_open(const char *path)
{
    depth_sensor *s = new depth_sensor();
    if (!s->open(path)) { delete s; return nullptr; }
    return s;
}
void kfh_sensor_close(void *h) { delete static_cast<depth_sensor *>(h); }
int kfh_sensor_info(void *h, float info8[8])
{
    const depth_sensor *s = static_cast<depth_sensor *>(h);
    info8[0] = (float)s->params.width; info8[1] = (float)s->params.height;
    info8[2] = s->params.fx; info8[3] = s->params.cx; info8[4] = s->params.fy; info8[5] = s->params.cy; info8[6] = s->params.c;
    info8[7] = (float)s->framesLeft();
    return 0;
}
/* next frame: depth_mm [h*w] float, bgr [h*w*3] (either may be NULL); 0 ok, 1 exhausted / undecodable, 2 size changed */
int kfh_sensor_get_frame(void *h, float *depth_mm, unsigned char *bgr)
{
    depth_sensor *s = static_cast<depth_sensor *>(h);
    if (!s->getFrame()) return 1;
    const size_t n = (size_t)s->params.width * s->params.height;
    if ((size_t)s->depth_map.rows * s->depth_map.cols != n || (size_t)s->color_map.rows * s->color_map.cols != n) return 2;
    if (depth_mm) std::memcpy(depth_mm, s->depth_map.ptr<float>(), n * sizeof(float));
    if (bgr) std::memcpy(bgr, s->color_map.ptr<unsigned char>(), n * 3);
    return 0;
}
const char *kfh_sensor_error(void *h) { return static_cast<depth_sensor *>(h)->lastError().c_str(); }
int kfh_png_write_gray16(const char *path, const unsigned short *pix, int w, int h) { return kf::png::write_gray16(path, pix, w, h) ? 0 : 1; }
int kfh_png_write_rgb8(const char *path, const unsigned char *rgb, int w, int h) { return kf::png::write_rgb8(path, rgb, w, h) ? 0 : 1; }
int kfh_icp_solve(const double in27[27], double x6[6]) { return kf::ICPRegistration::solve(in27, x6) ? 0 : 1; }

} // extern "C"
