// tsdf_volume.hpp -- kf::TSDFVolume: the map.  Public calls are those of the reference class
// (kfusion/include/tsdf_volume.hpp:10-46) minus the GpuMat arguments: the voxels and the frames both live behind the
// C-ABI context, so integrate / raycast only take the camera pose.
#pragma once
#include <string>
#include <vector>
#include "types.hpp"

namespace kf
{
class TSDFVolume
{
    DeviceContextPtr dev_;            // owner of the device-side volume
    cv::Vec3i grid_;                  // voxels per axis
    cv::Vec3f extent_, cell_;         // metres per axis: whole volume, one voxel
    cv::Affine3f pose_;               // volume -> world
    Intrinsics camera_{};
    float truncation_ = 0.f;          // metres
    int weight_cap_ = 64;

public:
    TSDFVolume() = default;
    // scene_size_: metres per axis, dims_: voxels per axis (tsdf_volume.cpp:14-19)
    TSDFVolume(const DeviceContextPtr &dev, const cv::Vec3f scene_size_, const cv::Vec3i dims_);
    ~TSDFVolume() { release(); }

    // running-average update from the context's current (filtered) depth frame; camera_pose = camera -> world.
    // Colour is dead state in the reference and is not integrated.
    void integrate(const cv::Affine3f &camera_pose);
    // model vertex / normal maps as seen from camera_pose, into the context's model frame
    void raycast(const cv::Affine3f &camera_pose);
    // zero-crossing points, 1 x N CV_32FC3 in the world frame
    cv::Mat fetchPointCloud();
    void reset();   // zero every voxel
    void release(); // drop the device context

    cv::Vec3i Dims() { return grid_; }
    cv::Vec3f SceneSize() { return extent_; }
    cv::Vec3f VoxelSize() { return cell_; }
    void setPose(const cv::Affine3f p) { pose_ = p; }
    void setIntrinsics(const Intrinsics i) { camera_ = i; }
    void setTrunDist(const float v) { truncation_ = v; }
    void setMaxWeight(const int w) { weight_cap_ = w; }

    // packed {tsdf, weight} int16 pairs, reference index order x + y X + z X Y
    std::vector<int16_t> Data();
    // volume checkpoint (no reference counterpart): header {magic, dims[3], range[3], trunc}
    // + the packed voxels in reference index order
    bool save(const std::string &path);
    bool load(const std::string &path);
};
} // namespace kf
