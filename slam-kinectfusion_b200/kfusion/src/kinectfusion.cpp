// kinectfusion.cpp -- the pipeline facade (mirrors kfusion/src/kinectfusion.cpp:9-196 of the
// reference: ctor wiring, imageProcess, the frame state machine, reset, render, PLY export,
// default parameters).  Device work is delegated to include/kfb200.h; this file holds only
// host orchestration and pose bookkeeping.
#include <kinectfusion.h>
#include <safe_call.hpp>
#include <fstream>
#include <iostream>
#include <stdexcept>

kf::kinectfusion::kinectfusion(const kf::Intrinsics intr, const kf::kinectfuison_params params) : vdata(nullptr), intr_(intr), params_(params)
{
    if ((int)params_.icp_iter_count.size() != params_.pyramid_height)
        throw std::invalid_argument("kf::kinectfusion: icp_iter_count needs one entry per pyramid level");
    kfb_params p;
    kfb_default_params(&p, params_.volu_dims(0));
    p.pyramid_height = params_.pyramid_height;
    p.dfilter_dist = params_.dfilter_dist;
    p.bfilter_kernel_size = params_.bfilter_kernel_size;
    p.bfilter_spatial_sigma = params_.bfilter_spatial_sigma;
    p.bfilter_color_sigma = params_.bfilter_color_sigma;
    p.icp_dist_threshold = params_.icp_dist_threshold;
    p.icp_angle_threshold = params_.icp_angle__threshold;
    for (int i = 0; i < KFB_MAX_LEVELS; ++i) p.icp_iter_count[i] = i < (int)params_.icp_iter_count.size() ? params_.icp_iter_count[i] : 0;
    for (int i = 0; i < 3; ++i) { p.volu_dims[i] = params_.volu_dims(i); p.volu_range[i] = params_.volu_range(i); }
    p.volu_trun_dist = params_.volu_trun_dist;
    p.tsdf_max_weight = params_.tsdf_max_weight;
    p.compat_icp_rows = params_.compat_icp_rows;
    p.compat_raycast_ts_sign = params_.compat_raycast_ts_sign;
    p.slab_z_begin = params_.slab_z_begin;
    p.slab_z_end = params_.slab_z_end;
    const kfb_intrinsics ki = intr_.abi();
    dev = std::make_shared<DeviceContext>();
    const int rc = kfb_create(&ki, &p, params_.device, &dev->ctx);
    if (rc != KFB_OK)
    {
        const std::string why = dev->ctx ? kfb_last_error_string(dev->ctx) : "invalid parameters";
        throw std::runtime_error("kf::kinectfusion: kfb_create failed: " + why);
    }
    //  -> handles onto the context's current / model frames
    cframe = Frame(dev, KFB_FRAME_CUR, params_.pyramid_height, intr_);
    pframe = Frame(dev, KFB_FRAME_PREV, params_.pyramid_height, intr_);
    frame_count = 1;
    vdata = new TSDFVolume(dev, params_.volu_range, params_.volu_dims);
    vdata->setMaxWeight(params_.tsdf_max_weight);
    vdata->setTrunDist(params_.volu_trun_dist);
    vdata->setIntrinsics(intr);
    vdata->setPose(params_.volu_pose);
    icp = ICPRegistration(params_.icp_dist_threshold, params_.icp_angle__threshold);
    icp.setIterationNum(params_.icp_iter_count);
    icp.setIntrinsics(intr);
    reset();
}
kf::kinectfusion::~kinectfusion() { release(); }

cv::Mat kf::kinectfusion::getRenderMap(DISPLAY_TYPES V)
{
    cv::Mat result(intr_.height, intr_.width, cv::CV_8UC3);
    if (V == NORMAL)
        kfbSafeCall(dev->ctx, kfb_render_normals(dev->ctx, result.ptr<uint8_t>()));
    else if (V == PHONG)
    {
        const cv::Vec3f t = pose_record.back().translation(); // kinectfusion.cpp:43
        kfbSafeCall(dev->ctx, kfb_render_phong(dev->ctx, t.val, result.ptr<uint8_t>()));
    }
    return result;
}

void kf::kinectfusion::imageProcess(const float *depth_mm, int width, int height)
{
    kfbCheck(dev->ctx, kfb_upload_depth_mm(dev->ctx, depth_mm, width, height));
    kfbCheck(dev->ctx, kfb_frontend(dev->ctx));
}

void kf::kinectfusion::pipeline(cv::Mat /*cmap_*/, cv::Mat dmap_)
{
    pipeline(dmap_.ptr<float>(), dmap_.cols, dmap_.rows);
}

void kf::kinectfusion::pipeline(const float *depth_mm, int width, int height)
{
    auto start_time = std::chrono::steady_clock::now();
    imageProcess(depth_mm, width, height);
    last_tracking_ok = true;
    if (frame_count == 1)
    {
        vdata->integrate(pose_record.back());
        kfbCheck(dev->ctx, kfb_swap_frames(dev->ctx)); // cframe->vmap.swap(pframe->vmap); nmap likewise (:88-89)
        kfbCheck(dev->ctx, kfb_upload_wait(dev->ctx));  // the caller may reuse its frame buffer when pipeline() returns
        frame_count++;
        return;
    }
    const bool sharded = params_.shard_world > 1;
    if (sharded && (!comm.broadcast_pose || !comm.composite)) throw std::runtime_error("kf::kinectfusion: sharded instance without ShardComm");
    // icp: transform of the current frame towards the previous one.  Sharded: ICP stays on rank 0 (it owns
    // the composited model maps) and only the pose travels.
    float msg[13] = {0.f};
    if (params_.shard_rank == 0)
    {
        cv::Affine3f cam_pose;
        const auto t_icp = std::chrono::steady_clock::now();
        const bool ok = icp.rigidTransform(cam_pose, pose_record.back(), &cframe, &pframe);
        last_icp_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_icp).count();
        msg[0] = ok ? 1.f : 0.f;
        if (ok) (pose_record.back() * cam_pose).to12(msg + 1);
    }
    if (sharded && comm.broadcast_pose(msg, comm.pose_user ? comm.pose_user : comm.user) != 0) throw std::runtime_error("kf::kinectfusion: pose broadcast failed");
    if (msg[0] == 0.f)
    {
        last_tracking_ok = false;
        std::cout << "tracking fail!" << std::endl;
        reset();
        return;
    }
    pose_record.push_back(cv::Affine3f::from12(msg + 1));
    vdata->integrate(pose_record.back());
    vdata->raycast(pose_record.back());
    if (sharded)
    {
        // first-hit composite over the slabs: one kernel over NVLink peer memory when the launcher attached the
        // peers (kfb_shard_attach), else the launcher's collectives
        if (kfb_shard_attached(dev->ctx)) kfbCheck(dev->ctx, kfb_shard_composite(dev->ctx));
        else if (comm.composite(comm.user) != 0) throw std::runtime_error("kf::kinectfusion: raycast composite failed");
    }
    if (params_.shard_rank == 0) kfbCheck(dev->ctx, kfb_model_pyramid(dev->ctx));
    else kfbCheck(dev->ctx, kfb_upload_wait(dev->ctx)); // rank 0 waited for the device in ICP; the others did not
    std::chrono::duration<double, std::milli> ms = std::chrono::steady_clock::now() - start_time;
    frame_time = std::to_string(ms.count());
    frame_count++;
}

cv::Affine3f kf::kinectfusion::getCurCameraPose()
{
    if (pose_record.size() > 0) return pose_record.back();
    return cv::Affine3f::Identity();
}
void kf::kinectfusion::reset()
{
    frame_count = 1;
    kfbSafeCall(dev->ctx, kfb_reset_frames(dev->ctx));
    vdata->reset();
    pose_record.clear();
    pose_record.push_back(cv::Affine3f::Identity());
}
cv::Mat kf::kinectfusion::extracePointcloud()
{
    points_array = vdata->fetchPointCloud();
    return points_array;
}
void kf::kinectfusion::savePointcloud(std::string path) { kf::file::exportPly(path, points_array); }

// kinectfusion.cpp:148-166: ASCII PLY, xyz only
void kf::file::exportPly(const std::string &filename, cv::Mat pointcloud)
{
    const int points_num = pointcloud.cols;
    std::ofstream file_out{filename};
    if (!file_out.is_open()) return;
    file_out << "ply" << std::endl;
    file_out << "format ascii 1.0" << std::endl;
    file_out << "element vertex " << points_num << std::endl;
    file_out << "property float x" << std::endl;
    file_out << "property float y" << std::endl;
    file_out << "property float z" << std::endl;
    file_out << "end_header" << std::endl;
    for (int i = 0; i < points_num; i++)
    {
        const float *p = pointcloud.ptr<float>() + 3 * (size_t)i;
        file_out << p[0] << " " << p[1] << " " << p[2] << "\n";
    }
}

// main.cpp:94-98
bool kf::file::exportPoses(const std::string &filename, const std::vector<cv::Affine3f> &poses)
{
    std::ofstream out{filename};
    if (!out.is_open()) return false;
    out.precision(8); // cv::Formatter default for float
    for (const cv::Affine3f &p : poses)
    {
        out << "[";
        for (int i = 0; i < 4; ++i)
        {
            for (int j = 0; j < 4; ++j) out << p.matrix(i, j) << (j < 3 ? ", " : "");
            out << (i < 3 ? ";\n " : "]");
        }
        out << std::endl;
    }
    return true;
}

// depth_sensor.cpp:23-46
bool kf::file::readIntrinsics(const std::string &filename, kf::Intrinsics &intr)
{
    std::ifstream in{filename};
    if (!in.is_open()) return false;
    std::vector<float> v;
    float t = 0;
    for (int i = 0; i < 9 && (in >> t); ++i)
        if (t > 0.1f) v.push_back(t);
    if (v.size() != 5) return false;
    intr.fx = v[0]; intr.cx = v[1]; intr.fy = v[2]; intr.cy = v[3]; intr.c = v[4];
    return true;
}

// The values of kinectfusion.cpp:167-190; the scalar ones are the struct's in-class initialisers.
kf::kinectfuison_params kf::kinectfuison_params::default_params()
{
    kf::kinectfuison_params p;
    p.volu_trun_dist = 2.1f * p.volu_range(0) / p.volu_dims(0);
    // the volume is centred on the first camera's optical axis and starts half a metre in front of it
    p.volu_pose = cv::Affine3f().translate(cv::Vec3f(-p.volu_range[0] / 2, -p.volu_range[1] / 2, 0.5f));
    return p;
}
void kf::kinectfusion::release()
{
    if (vdata) { delete vdata; vdata = nullptr; }
    cframe.dev.reset();
    pframe.dev.reset();
    dev.reset();
}

// ---- Frame accessors -------------------------------------------------------------------------------
cv::Mat kf::Frame::depth(int level) const
{
    const Intrinsics k = intr.level(level);
    cv::Mat m(k.height, k.width, cv::CV_32FC1);
    kfbSafeCall(dev->ctx, kfb_download_depth(dev->ctx, level, m.ptr<float>()));
    return m;
}
cv::Mat kf::Frame::vertices(int level) const
{
    const Intrinsics k = intr.level(level);
    cv::Mat m(k.height, k.width, cv::CV_32FC3);
    kfbSafeCall(dev->ctx, kfb_download_maps(dev->ctx, which, level, m.ptr<float>(), nullptr));
    return m;
}
cv::Mat kf::Frame::normals(int level) const
{
    const Intrinsics k = intr.level(level);
    cv::Mat m(k.height, k.width, cv::CV_32FC3);
    kfbSafeCall(dev->ctx, kfb_download_maps(dev->ctx, which, level, nullptr, m.ptr<float>()));
    return m;
}
