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
// Field names (and the misspelt type name) are the reference's (kinectfusion.h:9-31), so `params.volu_dims = ...`
// written against it keeps compiling; the in-class values are those of default_params() (kinectfusion.cpp:167-190).
struct kinectfuison_params
{
    // -- depth front end
    int pyramid_height = 3;                   // levels of the depth / vertex / normal pyramids
    float dfilter_dist = 5.f;                 // depth beyond this many metres is dropped
    int bfilter_kernel_size = 5;              // bilateral filter: kernel, spatial sigma (px), range sigma (mm)
    float bfilter_spatial_sigma = 10.f;
    float bfilter_color_sigma = 10.f;
    // -- tracking
    float icp_dist_threshold = 0.015f;        // correspondence gates: metres, degrees
    float icp_angle__threshold = 30.f;
    std::vector<int> icp_iter_count{4, 5, 10}; // iterations at level 0, 1, 2 (the coarse level runs first)
    // -- mapping: a volu_dims voxel grid over volu_range metres, placed by volu_pose
    cv::Vec3i volu_dims = cv::Vec3i::all(512);
    cv::Vec3f volu_range = cv::Vec3f::all(3.f);
    cv::Affine3f volu_pose;                   // default_params(): centred in x / y, 0.5 m in front of the first camera
    float volu_trun_dist = 2.1f * 3.f / 512;  // truncation distance: 2.1 voxels
    int tsdf_max_weight = 64;
    float init_cam_model_dist = 0.f;          // carried, never read (in the reference too)
    float min_pose_move = 0.008f;             // carried, never read (in the reference too)
    // -- this library's additions: bug-compat switches and placement
    int compat_icp_rows = 1;
    int compat_raycast_ts_sign = 1;
    int device = 0;
    // -- z-slab sharding (volumes of 1024^3 and above): this instance stores and updates planes
    //    [slab_z_begin, slab_z_end) (+ halo) of the volume; 0,0 = the whole volume
    int slab_z_begin = 0, slab_z_end = 0;
    int shard_rank = 0, shard_world = 1;

    static kinectfuison_params default_params();
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
    void *pose_user = nullptr; // argument of broadcast_pose when it differs from `user` (kf::PoseMailbox::callback)
};

// The tracker-and-mapper.  Construction allocates everything on the device (one kfb_ctx); pipeline() is the only
// per-frame call.  Public names follow kfusion/include/kinectfusion.h:32-73 of the reference.
class kinectfusion
{
    DeviceContextPtr dev;            // the C-ABI context: volume, frame pyramids, ICP buffers
    Frame cframe, pframe;            // handles on the context's current frame / model (raycast) frame
    TSDFVolume *vdata = nullptr;
    ICPRegistration icp;
    Intrinsics intr_;
    kinectfuison_params params_;
    ShardComm comm;                  // collectives of a sharded instance (unset: single GPU)
    cv::Mat points_array;            // last extracted cloud, 1 x N CV_32FC3
    void imageProcess(const float *depth_mm, int width, int height);

public:
    enum DISPLAY_TYPES { PHONG, NORMAL };

    kinectfusion(const kf::Intrinsics intr, const kf::kinectfuison_params params);
    ~kinectfusion();

    // One frame: ingest + front end, ICP against the model frame (from the second frame on), integrate, raycast.
    // dmap_: CV_32FC1 depth in millimetres; cmap_: CV_8UC3 BGR, accepted and ignored (write-only state in the
    // reference).  On a tracking failure prints "tracking fail!" and resets, like kinectfusion.cpp:97-102.
    void pipeline(cv::Mat cmap_, cv::Mat dmap_);
    // the same from a raw (ideally pinned) host pointer: no copy on the host side
    void pipeline(const float *depth_mm, int width, int height);
    // back to the state after construction: empty volume, identity pose, frame counter 1
    void reset();
    void release();

    cv::Affine3f getCurCameraPose();                 // camera -> world of the last tracked frame
    cv::Mat getRenderMap(DISPLAY_TYPES V = PHONG);   // CV_8UC3 view of the model frame
    cv::Mat extracePointcloud();                     // zero-crossing points of the volume, world frame
    void savePointcloud(std::string path);           // the last extracted cloud as ASCII PLY

    // this library's additions
    void setShardComm(const ShardComm &c) { comm = c; }
    kfb_ctx *context() { return dev ? dev->ctx : nullptr; }
    TSDFVolume *volume() { return vdata; }
    const Frame *currentFrame() const { return &cframe; }
    const Frame *modelFrame() const { return &pframe; }

    // state the reference exposes as public members
    int frame_count = 1;                      // 1 before the first frame
    std::string frame_time;                   // milliseconds of the last pipeline() call, as text
    std::vector<cv::Affine3f> pose_record;    // one camera pose per processed frame
    bool last_tracking_ok = true;
    double last_icp_us = 0.0;                 // wall time of the last ICPRegistration::rigidTransform (diagnostic)
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
