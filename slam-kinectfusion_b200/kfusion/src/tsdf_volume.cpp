// tsdf_volume.cpp -- kf::TSDFVolume host side: pose algebra and hand-off to the C-ABI
// (mirrors kfusion/src/tsdf_volume.cpp:13-84 of the reference).
#include <tsdf_volume.hpp>
#include <safe_call.hpp>
#include <cstdio>
#include <cstring>

namespace kf
{
std::vector<int16_t> TSDFVolume::Data()
{
    std::vector<int16_t> out(2 * kfb_volume_voxels(dev_->ctx));
    kfbSafeCall(dev_->ctx, kfb_download_volume(dev_->ctx, out.data()));
    return out;
}
namespace
{
struct CheckpointHeader
{
    char magic[8]; // "KFB200V1"
    int32_t dims[3];
    float range[3];
    float trunc;
    uint64_t voxels;
};
} // namespace
bool TSDFVolume::save(const std::string &path)
{
    const std::vector<int16_t> data = Data();
    CheckpointHeader h;
    std::memcpy(h.magic, "KFB200V1", 8);
    for (int i = 0; i < 3; ++i) { h.dims[i] = grid_(i); h.range[i] = extent_(i); }
    h.trunc = truncation_;
    h.voxels = data.size() / 2;
    std::FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(data.data(), sizeof(int16_t), data.size(), f) == data.size();
    ok = (std::fclose(f) == 0) && ok;
    return ok;
}
bool TSDFVolume::load(const std::string &path)
{
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    CheckpointHeader h;
    bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "KFB200V1", 8) == 0;
    for (int i = 0; i < 3 && ok; ++i) ok = h.dims[i] == grid_(i) && h.range[i] == extent_(i);
    ok = ok && h.voxels == kfb_volume_voxels(dev_->ctx);
    std::vector<int16_t> data;
    if (ok)
    {
        data.resize(2 * (size_t)h.voxels);
        ok = std::fread(data.data(), sizeof(int16_t), data.size(), f) == data.size();
    }
    std::fclose(f);
    if (!ok) return false;
    return kfbSafeCall(dev_->ctx, kfb_upload_volume(dev_->ctx, data.data())) == KFB_OK;
}

TSDFVolume::TSDFVolume(const DeviceContextPtr &dev_arg, const cv::Vec3f scene_size_, const cv::Vec3i dims_)
    : dev_(dev_arg), grid_(dims_), extent_(scene_size_)
{
    cell_ = cv::Vec3f(scene_size_(0) / dims_(0), scene_size_(1) / dims_(1), scene_size_(2) / dims_(2)); // :16
    reset();
}
void TSDFVolume::reset() { if (dev_) kfbSafeCall(dev_->ctx, kfb_reset_volume(dev_->ctx)); }
void TSDFVolume::release() { dev_.reset(); }

void TSDFVolume::integrate(const cv::Affine3f &camera_pose)
{
    const cv::Affine3f vol2cam = camera_pose.inv() * pose_; // :50
    float p[12];
    vol2cam.to12(p);
    kfbCheck(dev_->ctx, kfb_integrate(dev_->ctx, p, nullptr));
}
void TSDFVolume::raycast(const cv::Affine3f &camera_pose)
{
    const cv::Affine3f cam2vol = pose_.inv() * camera_pose; // :59
    // cam2vol.rotation().inv(DECOMP_SVD) (:61): inverse of the rotation block
    cv::Affine3f rot_only(cam2vol.rotation(), cv::Vec3f(0.f, 0.f, 0.f));
    const cv::Matx33f Rinv = rot_only.inv().rotation();
    float p[12];
    cam2vol.to12(p);
    kfbCheck(dev_->ctx, kfb_raycast(dev_->ctx, p, Rinv.val));
}
cv::Mat TSDFVolume::fetchPointCloud()
{
    enum { DEFAULT_CLOUD_BUFFER_SIZE = 10 * 1000 * 1000 }; // :65-68
    std::vector<float> pts((size_t)DEFAULT_CLOUD_BUFFER_SIZE * 3);
    float vp[12];
    pose_.to12(vp);
    size_t n = 0;
    kfbSafeCall(dev_->ctx, kfb_extract_points(dev_->ctx, vp, pts.data(), DEFAULT_CLOUD_BUFFER_SIZE, &n));
    return cv::Mat(1, (int)n, cv::CV_32FC3, pts.data());
}
} // namespace kf
