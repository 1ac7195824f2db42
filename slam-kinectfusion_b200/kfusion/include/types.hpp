// types.hpp -- kf::Intrinsics and kf::Frame (mirrors kfusion/include/types.hpp:13-80 of the reference).
#pragma once
#include <cmath>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
#include "cvlite.hpp"
#include "../../../include/kfb200.h"

#define STR(v) #v

namespace kf
{
// types.hpp:13-29
struct Intrinsics
{
    int width, height;
    float fx, fy, cx, cy;
    float c = 1; // depth scale, carried but unused exactly as in the reference (image_process.cu:14 hard-codes 0.001)
    Intrinsics level(const size_t level) const
    {
        if (level == 0) return *this;
        const float scale_factor = powf(0.5f, static_cast<float>(level));
        return Intrinsics{width >> level, height >> level, fx * scale_factor, fy * scale_factor,
                          (cx + 0.5f) * scale_factor - 0.5f, (cy + 0.5f) * scale_factor - 0.5f};
    }
    kfb_intrinsics abi() const { return kfb_intrinsics{width, height, fx, fy, cx, cy}; }
};

// Owner of the C-ABI context (all device memory lives behind it).
struct DeviceContext
{
    kfb_ctx *ctx = nullptr;
    ~DeviceContext() { if (ctx) kfb_destroy(ctx); }
};
typedef std::shared_ptr<DeviceContext> DeviceContextPtr;

// types.hpp:30-80.  The reference's Frame owns per-level GpuMats; here the maps live in the
// context and a Frame is a handle (current or model frame) with the download accessors tests need.
struct Frame
{
    DeviceContextPtr dev;
    int which;       // KFB_FRAME_CUR / KFB_FRAME_PREV
    int pyr_height;
    Intrinsics intr;
    Frame() : which(KFB_FRAME_CUR), pyr_height(0), intr{} {}
    Frame(const DeviceContextPtr &d, int which_, int pyr_height_, const Intrinsics &i) : dev(d), which(which_), pyr_height(pyr_height_), intr(i) {}
    cv::Mat depth(int level = 0) const;   // CV_32FC1 metres (current frame only)
    cv::Mat vertices(int level = 0) const; // CV_32FC3
    cv::Mat normals(int level = 0) const;  // CV_32FC3
};
inline float deg2rad(float alpha) { return alpha * 0.017453293f; }
} // namespace kf
