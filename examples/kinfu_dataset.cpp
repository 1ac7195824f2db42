// kinfu_dataset.cpp -- the reference application's main loop (main.cpp:24-134 of baiyuntao00/SLAM-KinectFusion),
// headless: same classes, same calls, same order; the cv::viz / cv::imshow windows and the keyboard handler are
// left out (out of scope), the Phong view of the last frame is written as a PNG instead.
//
//   kinfu_dataset <dataset dir> [out dir]      dataset = color/*.png, depth/*.png (16-bit mm), intr.txt
//
// Build: make -C slam-kinectfusion_b200/kfusion example
#include <depth_sensor.h>
#include <kinectfusion.h>

#include <fstream>
#include <iostream>
#include <string>

struct KinectFusionAPP
{
    depth_sensor *camera;
    kf::kinectfuison_params kfparams;
    kf::kinectfusion *kinfu;
    std::string out;

    KinectFusionAPP(depth_sensor *camera_, const std::string &out_) : camera(camera_), out(out_)
    {
        kfparams = kfparams.default_params();
        kinfu = new kf::kinectfusion(camera->params, kfparams);
    }
    bool execute()
    {
        cv::Mat scene;
        for (;;)
        {
            if (!camera->getFrame())
            {
                std::cout << "no image!" << std::endl;
                break;
            }
            kinfu->pipeline(camera->color_map, camera->depth_map);
            scene = kinfu->getRenderMap(kinfu->PHONG);
            if (kinfu->frame_count % 5 == 0) kinfu->extracePointcloud(); // the reference refreshes its 3-D view here
        }
        if (!scene.empty())
        {
            // BGR -> RGB for the PNG writer
            cv::Mat rgb(scene.rows, scene.cols, cv::CV_8UC3);
            const unsigned char *s = scene.ptr<unsigned char>();
            unsigned char *d = rgb.ptr<unsigned char>();
            for (size_t i = 0; i < (size_t)scene.rows * scene.cols; ++i) { d[3 * i] = s[3 * i + 2]; d[3 * i + 1] = s[3 * i + 1]; d[3 * i + 2] = s[3 * i]; }
            kf::png::write_rgb8(out + "/scene.png", d, scene.cols, scene.rows);
        }
        kinfu->extracePointcloud();
        kinfu->savePointcloud(out + "/pointcloud.ply");
        // output camera poses (main.cpp:95-98)
        kf::file::exportPoses(out + "/poses.txt", kinfu->pose_record);
        std::cout << kinfu->pose_record.size() << " frames, end!" << std::endl;
        return true;
    }
    void release()
    {
        camera->release();
        kinfu->release();
    }
};

int main(int argc, char *argv[])
{
    std::cout << "KinectFusion: start" << std::endl;
    depth_sensor camera;
    if (!camera.open(argc > 1 ? argv[1] : "../../dataset"))
    {
        std::cout << camera.lastError() << std::endl;
        return 1;
    }
    KinectFusionAPP app(&camera, argc > 2 ? argv[2] : ".");
    try
    {
        app.execute();
        app.release();
    }
    catch (const std::bad_alloc &)
    {
        std::cout << "Bad alloc" << std::endl;
    }
    catch (const std::exception &)
    {
        std::cout << "Exception" << std::endl;
    }
    return 0;
}
