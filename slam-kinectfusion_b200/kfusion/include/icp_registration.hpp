// icp_registration.hpp -- kf::ICPRegistration (mirrors kfusion/include/icp_registration.hpp:7-26).
#pragma once
#include <vector>
#include "types.hpp"

namespace kf
{
class ICPRegistration
{
public:
    ICPRegistration() {}
    ICPRegistration(const float d, const float a);
    ~ICPRegistration() {}

    void setMaxDistThres(const float max_dist_);
    void setMaxAngleThres(const float max_angle_);
    void setIterationNum(const std::vector<int> &iters);
    void setIntrinsics(const Intrinsics intrs_);

    // Relative transform current -> previous; false => tracking failure (caller resets).
    bool rigidTransform(cv::Affine3f &curpose, const cv::Affine3f prepose, const Frame *cframe, const Frame *pframe);

    // 6x6 normal-equation solve with the reference's determinant guard (icp_registration.cpp:35-39);
    // Cholesky in double (LU fallback).  Returns false on |det| < 1e-15 or NaN.
    static bool solve(const double sums27[27], double x6[6]);

private:
    std::vector<int> iters;
    Intrinsics intrs;
    float angle_thres;
    float dist_thres;
};
} // namespace kf
