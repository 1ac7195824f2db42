// depth_sensor.h -- the DATASET flavour of the reference's frame source (kfusion/include/depth_sensor.h:36-50,
// kfusion/src/depth_sensor.cpp:11-46,186-196): a directory with color/*.png, depth/*.png (16-bit millimetres) and
// intr.txt.  Same public members and calls as the reference class, so main.cpp's loop
//     depth_sensor camera(path); while (camera.getFrame()) kinfu.pipeline(camera.color_map, camera.depth_map);
// compiles against it.  The Kinect 2 / RealSense SDK flavours are out of scope (no SDKs here).  PNG files are read
// by a small decoder of our own on zlib (src/depth_sensor.cpp): the reference uses cv::imread, and OpenCV's C++ library
// is not part of this build.
#pragma once
#include <algorithm>
#include <string>
#include <vector>
#include "types.hpp"

namespace kf
{
namespace png
{
// Decoded image: 8- or 16-bit samples, host byte order, `channels` interleaved samples per pixel as stored in the
// file (1 grey, 2 grey+alpha, 3 RGB, 4 RGBA; palette images are expanded to RGB).
struct Image
{
    int width = 0, height = 0, channels = 0, bit_depth = 0;
    std::vector<unsigned char> data8;   // bit_depth <= 8 (1/2/4-bit samples are widened to 8 without scaling)
    std::vector<unsigned short> data16; // bit_depth == 16
};
// false + `err` on anything that is not a well-formed PNG (all colour types and bit depths, Adam7 interlace)
bool read(const std::string &path, Image &out, std::string *err = nullptr);
// writer for tests and tools: 8-bit RGB (channels == 3) or 16-bit grey (channels == 1)
bool write_gray16(const std::string &path, const unsigned short *pix, int width, int height);
bool write_rgb8(const std::string &path, const unsigned char *rgb, int width, int height);
} // namespace png
} // namespace kf

class depth_sensor
{
public:
    kf::Intrinsics params{640, 480, 0.f, 0.f, 0.f, 0.f};
    cv::Mat color_map; // CV_8UC3, BGR like cv::imread(name, 1) (depth_sensor.cpp:189)
    cv::Mat depth_map; // CV_32FC1, the file's 16-bit values converted to float (depth_sensor.cpp:191)
    const int width = 640;  // the reference's compile-time defaults (depth_sensor.h:37-38); `params` holds the
    const int height = 480; // real size once a dataset is open

    depth_sensor() {}
    depth_sensor(const std::string &path) { open(path); }
    ~depth_sensor() { release(); }
    // false: no color/*.png or depth/*.png under `path` (the reference prints "error: no camera!" and exits)
    bool open(const std::string &path);
    // next frame into color_map / depth_map; false when the list is exhausted or a file cannot be decoded
    bool getFrame();
    void release();
    size_t framesLeft() const { return std::min(img_col_name.size(), img_dep_name.size()); }
    const std::string &lastError() const { return err; }

private:
    std::string data_path, err;
    std::vector<std::string> img_col_name, img_dep_name; // sorted by name (cv::glob sorts)
};
