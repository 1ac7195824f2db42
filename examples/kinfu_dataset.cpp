// kinfu_dataset.cpp -- headless dataset runner written against the reference's public classes as this repo
// provides them (depth_sensor, kf::kinectfusion, kf::kinectfuison_params): what an application built on
// baiyuntao00/SLAM-KinectFusion does per frame (main.cpp:64-101: grab, pipeline, render, periodically extract)
// and at exit (poses.txt, main.cpp:95-98), without the cv::viz / cv::imshow windows (out of scope).
//
//   kinfu_dataset <dataset dir> [out dir] [volume side, default 512]
//   dataset = color/*.png, depth/*.png (16-bit millimetres), intr.txt
//
// Build: make -C slam-kinectfusion_b200/kfusion example
#include <depth_sensor.h>
#include <kinectfusion.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

namespace
{
// the Phong view comes back BGR (like the reference's cv::Mat); the PNG writer wants RGB
bool save_view(const std::string &path, const cv::Mat &bgr)
{
    if (bgr.empty()) return false;
    cv::Mat rgb(bgr.rows, bgr.cols, cv::CV_8UC3);
    const unsigned char *src = bgr.ptr<unsigned char>();
    unsigned char *dst = rgb.ptr<unsigned char>();
    for (size_t i = 0, n = (size_t)bgr.rows * bgr.cols; i < n; ++i)
    {
        dst[3 * i + 0] = src[3 * i + 2];
        dst[3 * i + 1] = src[3 * i + 1];
        dst[3 * i + 2] = src[3 * i + 0];
    }
    return kf::png::write_rgb8(path, dst, bgr.cols, bgr.rows);
}

int run(const std::string &dataset, const std::string &out_dir, int side)
{
    depth_sensor source;
    if (!source.open(dataset))
    {
        std::fprintf(stderr, "%s: %s\n", dataset.c_str(), source.lastError().c_str());
        return 1;
    }
    if (!(source.params.fx > 0.f))
    {
        std::fprintf(stderr, "%s/intr.txt is missing or malformed (fx 0 cx / 0 fy cy / 0 0 1)\n", dataset.c_str());
        return 1;
    }
    kf::kinectfuison_params params = kf::kinectfuison_params::default_params();
    params.volu_dims = cv::Vec3i::all(side);
    params.volu_trun_dist = 2.1f * params.volu_range(0) / params.volu_dims(0);
    kf::kinectfusion fusion(source.params, params);

    const size_t total = source.framesLeft();
    size_t lost = 0;
    cv::Mat view;
    const auto t0 = std::chrono::steady_clock::now();
    while (source.getFrame())
    {
        fusion.pipeline(source.color_map, source.depth_map);
        lost += fusion.last_tracking_ok ? 0 : 1;
        view = fusion.getRenderMap(kf::kinectfusion::PHONG);
        if (fusion.pose_record.size() % 5 == 0) fusion.extracePointcloud(); // the reference refreshes its 3-D view every 5th frame
    }
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    const size_t done = fusion.pose_record.size();
    if (done != total) std::fprintf(stderr, "stopped after %zu of %zu frames: %s\n", done, total, source.lastError().c_str());

    save_view(out_dir + "/scene.png", view);
    fusion.extracePointcloud();
    fusion.savePointcloud(out_dir + "/pointcloud.ply");
    kf::file::exportPoses(out_dir + "/poses.txt", fusion.pose_record);
    std::printf("%zu frames (%zu tracking failures), %.2f ms per frame including PNG decoding -> %s/{poses.txt,pointcloud.ply,scene.png}\n",
                done, lost, done ? ms / done : 0.0, out_dir.c_str());
    fusion.release();
    return done == total ? 0 : 2;
}
} // namespace

int main(int argc, char **argv)
{
    if (argc < 2)
    {
        std::fprintf(stderr, "usage: %s <dataset dir> [out dir] [volume side]\n", argv[0]);
        return 64;
    }
    return run(argv[1], argc > 2 ? argv[2] : ".", argc > 3 ? std::atoi(argv[3]) : 512);
}
