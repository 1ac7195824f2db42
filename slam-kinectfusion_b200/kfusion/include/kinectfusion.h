// kinectfusion.h -- kf::kinectfusion and kf::kinectfuison_params (sic), mirroring
// kfusion/include/kinectfusion.h:9-73 of the reference name for name (misspellings kept so
// existing call sites port mechanically).  All device work goes through include/kfb200.h.
#pragma once
#include <chrono>
#include <string>
#include <vector>
#include "types.hpp"
#include "tsdf_volume.hpp"
#include "icp_registration.hpp"

namespace kf
{
struct kinectfuison_params
{
    static kinectfuison_params default_params();
    ////surf meaasure
    int pyramid_height;
    float dfilter_dist;
    int bfilter_kernel_size;
    float bfilter_spatial_sigma;
    float bfilter_color_sigma;
    ////pose estimation
    float icp_dist_threshold;
    float icp_angle__threshold;
    std::vector<int> icp_iter_count;
    ////volume fusion
    cv::Vec3f volu_range;
    cv::Affine3f volu_pose;
    float volu_trun_dist;
    float init_cam_model_dist; // dead in the reference too
    cv::Vec3i volu_dims;
    float min_pose_move;       // dead in the reference too
    int tsdf_max_weight;
    //// additions (bug-compat switches) and placement
    int compat_icp_rows = 1;
    int compat_raycast_ts_sign = 1;
    int device = 0;
    //// z-slab sharding (volumes of 1024^3 and above): this instance stores and updates planes
    //// [slab_z_begin, slab_z_end) (+ halo) of the volume; 0,0 = the whole volume
    int slab_z_begin = 0, slab_z_end = 0;
    int shard_rank = 0, shard_world = 1;
};

// Collectives a sharded instance needs, supplied by the launcher (one process per GPU; the bench harness
// implements them with torch.distributed / NCCL on the context's stream).  Both return 0 on success.
struct ShardComm
{
    // msg13 = {tracking_ok, pose12 of the new global camera pose}: valid on rank 0 on entry, on every rank on return
    int (*broadcast_pose)(float *msg13, void *user) = nullptr;
    // cross-slab raycast composite after kfb_raycast: all-reduce MIN of the event keys, kfb_composite_mask,
    // integer SUM reduction of the model maps to rank 0 (include/kfb200.h, kfb_composite_mask)
    int (*composite)(void *user) = nullptr;
    void *user = nullptr;
};

class kinectfusion
{
public:
    kinectfusion(const kf::Intrinsics intr, const kf::kinectfuison_params params);
    ~kinectfusion();

    // dmap_: CV_32FC1 depth in millimetres; cmap_: CV_8UC3 BGR (accepted, write-only state in the reference)
    void pipeline(cv::Mat cmap_, cv::Mat dmap_);
    // same, from a raw (ideally pinned) host pointer: no copy on the host side
    void pipeline(const float *depth_mm, int width, int height);
    void reset();
    enum DISPLAY_TYPES
    {
        PHONG,
        NORMAL,
    };
    cv::Mat getRenderMap(DISPLAY_TYPES V = PHONG);
    cv::Mat extracePointcloud();
    void savePointcloud(std::string path);
    cv::Affine3f getCurCameraPose();
    void release();

    void setShardComm(const ShardComm &c) { comm = c; }
    kfb_ctx *context() { return dev ? dev->ctx : nullptr; }
    const Frame *currentFrame() const { return &cframe; }
    const Frame *modelFrame() const { return &pframe; }
    TSDFVolume *volume() { return vdata; }

public:
    std::string frame_time;
    int frame_count;
    std::vector<cv::Affine3f> pose_record;
    bool last_tracking_ok = true;
    double last_icp_us = 0.0; // wall time of the last ICPRegistration::rigidTransform (diagnostic)

private:
    void imageProcess(const float *depth_mm, int width, int height);

private:
    DeviceContextPtr dev;
    Frame cframe;
    Frame pframe;
    TSDFVolume *vdata;
    ICPRegistration icp;
    Intrinsics intr_;
    kinectfuison_params params_;
    ShardComm comm;
    cv::Mat points_array;
};
} // namespace kf

namespace kf
{
namespace file
{
void exportPly(const std::string &filename, cv::Mat pointcloud);
// camera trajectory in the reference's format (main.cpp:94-98: `outfile << pose.matrix << std::endl`, the
// cv::Matx stream form "[a, b, c, d;\n e, ...]"), one 4x4 per pose
bool exportPoses(const std::string &filename, const std::vector<cv::Affine3f> &poses);
// the dataset's intr.txt (depth_sensor.cpp:23-46): up to nine numbers, those > 0.1 are fx, cx, fy, cy, depth scale
bool readIntrinsics(const std::string &filename, kf::Intrinsics &intr);
}
}
