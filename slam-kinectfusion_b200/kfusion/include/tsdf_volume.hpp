// tsdf_volume.hpp -- kf::TSDFVolume (mirrors kfusion/include/tsdf_volume.hpp:10-46).
#pragma once
#include "types.hpp"

namespace kf
{
class TSDFVolume
{
public:
    TSDFVolume() {}
    ~TSDFVolume() { release(); }
    TSDFVolume(const DeviceContextPtr &dev, const cv::Vec3f scene_size_, const cv::Vec3i dims_);

    void setTrunDist(const float v);
    void setMaxWeight(const int w);
    void setPose(const cv::Affine3f p);
    void setIntrinsics(const Intrinsics i);

    void release();
    void reset();

    // dmap/cmap of the reference live in the context (current frame); colour is dead state
    void integrate(const cv::Affine3f &camera_pose);
    void raycast(const cv::Affine3f &camera_pose);
    cv::Mat fetchPointCloud(); // 1 x N CV_32FC3, world frame

    cv::Vec3f VoxelSize();
    cv::Vec3f SceneSize();
    cv::Vec3i Dims();
    std::vector<int16_t> Data(); // packed {tsdf, weight} pairs, reference index order
    // volume checkpoint (no reference counterpart): header {magic, dims[3], range[3], trunc}
    // + the packed voxels in reference index order
    bool save(const std::string &path);
    bool load(const std::string &path);

private:
    DeviceContextPtr dev;
    cv::Affine3f volume_pose;
    cv::Vec3f scene_size;
    cv::Vec3f voxel_size;
    cv::Vec3i dims;
    Intrinsics intr;
    float trun_dist = 0.f;
    int max_weight = 64;
};
} // namespace kf
