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

    kfb_ctx *context() { return dev ? dev->ctx : nullptr; }
    const Frame *currentFrame() const { return &cframe; }
    const Frame *modelFrame() const { return &pframe; }
    TSDFVolume *volume() { return vdata; }

public:
    std::string frame_time;
    int frame_count;
    std::vector<cv::Affine3f> pose_record;
    bool last_tracking_ok = true;

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
    cv::Mat points_array;
};
} // namespace kf

namespace kf
{
namespace file
{
void exportPly(const std::string &filename, cv::Mat pointcloud);
}
}
