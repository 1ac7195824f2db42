// icp_registration.hpp -- kf::ICPRegistration: the host side of the tracking step.  Public calls and their meaning
// are those of the reference class (kfusion/include/icp_registration.hpp:7-26, src/icp_registration.cpp), so code
// written against it compiles unchanged; everything behind them is this library's.
#pragma once
#include <vector>
#include "types.hpp"

namespace kf
{
class ICPRegistration
{
    // gates as the device expects them: metres, and the SINE of the angle (icp_registration.cpp:5)
    float gate_distance_ = 0.f, gate_sine_ = 0.f;
    // iterations per pyramid level, index = level (coarse levels last in the vector, run first)
    std::vector<int> schedule_;
    Intrinsics camera_{};

public:
    ICPRegistration() = default;
    // d: distance gate in metres, a: angle gate in degrees
    ICPRegistration(const float d, const float a);

    // Estimates the camera pose of the current frame against the model frame, coarse to fine, starting from the
    // previous pose.  Writes `curpose`; false = tracking failure (singular system), the caller resets.
    bool rigidTransform(cv::Affine3f &curpose, const cv::Affine3f prepose, const Frame *cframe, const Frame *pframe);

    void setIterationNum(const std::vector<int> &per_level) { schedule_ = per_level; }
    void setIntrinsics(const Intrinsics k) { camera_ = k; }
    void setMaxDistThres(const float metres) { gate_distance_ = metres; }
    // stores the angle in radians, not its sine -- the reference's setter does the same (icp_registration.cpp:10)
    void setMaxAngleThres(const float degrees) { gate_sine_ = deg2rad(degrees); }

    // The 6x6 normal equations from the 27 packed sums: determinant guard of the reference
    // (icp_registration.cpp:35-39: |det| < 1e-15 or NaN => false), then square-root-free Cholesky in double with
    // an LU fallback.  Operation order is mirrored on the device (kfb_icp.cu: icp_predict_pose).
    static bool solve(const double sums27[27], double x6[6]);
};
} // namespace kf
