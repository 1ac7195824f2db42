// icp_registration.cpp -- coarse-to-fine ICP loop and the host 6x6 solve
// (mirrors kfusion/src/icp_registration.cpp:3-45 of the reference).
#include <icp_registration.hpp>
#include <safe_call.hpp>
#include <cstring>

kf::ICPRegistration::ICPRegistration(const float d, const float a) : gate_distance_(d), gate_sine_(sinf(deg2rad(a))) {}

// rigid_icp.cu:156-165 (unpack) + icp_registration.cpp:35-39 (guard + solve)
bool kf::ICPRegistration::solve(const double in27[27], double x6[6])
{
    double A[6][6], b[6];
    int s = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 7; ++j)
        {
            const double v = in27[s++];
            if (j == 6) b[i] = v;
            else A[i][j] = A[j][i] = v;
        }
    // determinant by LU with partial pivoting (cv::determinant on a Matx66d)
    double M[6][7];
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < 6; ++j) M[i][j] = A[i][j]; M[i][6] = b[i]; }
    double det = 1.0;
    bool singular = false;
    for (int c = 0; c < 6; ++c)
    {
        int p = c;
        for (int r = c + 1; r < 6; ++r) if (std::fabs(M[r][c]) > std::fabs(M[p][c])) p = r;
        if (M[p][c] == 0.0 || std::isnan(M[p][c])) { det = std::isnan(M[p][c]) ? NAN : 0.0; singular = true; break; }
        if (p != c) { for (int j = 0; j < 7; ++j) std::swap(M[c][j], M[p][j]); det = -det; }
        det *= M[c][c];
        for (int r = c + 1; r < 6; ++r)
        {
            const double f = M[r][c] / M[c][c];
            for (int j = c; j < 7; ++j) M[r][j] -= f * M[c][j];
        }
    }
    if (singular || std::fabs(det) < 1e-15 || std::isnan(det)) return false;
    // Cholesky in its square-root-free form A = L D L^T with one reciprocal per column (north star: host 6x6
    // Cholesky solve); LU back-substitution if A is not numerically positive definite.  The device computes
    // the next pose with EXACTLY this operation sequence (csrc/kfb_icp.cu: icp_predict_pose) -- keep the two in
    // step: every product, difference and reciprocal below is one IEEE double operation (-ffp-contract=off).
    double L[6][6], d[6], inv[6], t[6][6];
    bool ok = true;
    for (int j = 0; j < 6 && ok; ++j)
    {
        double dj = A[j][j];
        for (int q = 0; q < j; ++q) { t[j][q] = L[j][q] * d[q]; dj -= L[j][q] * t[j][q]; }
        if (!(dj > 0.0)) { ok = false; break; }
        d[j] = dj;
        inv[j] = 1.0 / dj;
        for (int i = j + 1; i < 6; ++i)
        {
            double sum = A[i][j];
            for (int q = 0; q < j; ++q) sum -= L[i][q] * t[j][q];
            L[i][j] = sum * inv[j];
        }
    }
    if (ok)
    {
        double z[6];
        for (int i = 0; i < 6; ++i) { double sum = b[i]; for (int q = 0; q < i; ++q) sum -= L[i][q] * z[q]; z[i] = sum; }
        for (int i = 5; i >= 0; --i) { double sum = z[i] * inv[i]; for (int q = i + 1; q < 6; ++q) sum -= L[q][i] * x6[q]; x6[i] = sum; }
    }
    else
    {
        for (int i = 5; i >= 0; --i) { double sum = M[i][6]; for (int q = i + 1; q < 6; ++q) sum -= M[i][q] * x6[q]; x6[i] = sum / M[i][i]; }
    }
    return true;
}

bool kf::ICPRegistration::rigidTransform(cv::Affine3f &camera_pose, const cv::Affine3f /*prepose*/, const Frame *cframe, const Frame * /*pframe*/)
{
    // The reference's `camera_pose.Identity()` (:18) is a no-op on a default-constructed pose; same start here.
    camera_pose = cv::Affine3f::Identity();
    kfb_ctx *ctx = cframe->dev->ctx;
    // The loop below is the reference's (:21-43).  On the device ONE kernel runs the whole schedule by itself,
    // computing every next pose with solve()'s arithmetic; kfb_icp_step hands over iteration k's sums once the pose
    // the kernel used for it has been found equal, bit for bit, to the pose passed in, so this loop stays the
    // authority on every pose while no launch, PCIe read or host solve sits between two accumulations.
    int sched[KFB_MAX_LEVELS] = {0};
    for (size_t l = 0; l < schedule_.size() && l < KFB_MAX_LEVELS; ++l) sched[l] = schedule_[l];
    // false is reserved for the reference's meaning (singular system => tracking failure => the caller resets the
    // map); a device or transport error is raised as kf::DeviceError instead and leaves the map alone
    kfbCheck(ctx, kfb_icp_begin(ctx, sched));
    bool ok = true;
    for (int level = (int)schedule_.size() - 1; level >= 0 && ok; level--)
    {
        for (int i = 0; i < schedule_[level] && ok; i++)
        {
            float pose12[12];
            double sums[27], x[6];
            camera_pose.to12(pose12);
            const int rc = kfb_icp_step(ctx, pose12, sums);
            if (rc != KFB_OK)
            {
                kfb_icp_end(ctx);
                kfbCheck(ctx, rc);
            }
            if (!solve(sums, x)) { ok = false; break; }
            // Tinc = Affine3f(rvec = x[0..2] (float), t = x[3..5]); pose = pose * Tinc (right-multiply, :41-42)
            cv::Affine3f Tinc(cv::Vec3f((float)x[0], (float)x[1], (float)x[2]), cv::Vec3f((float)x[3], (float)x[4], (float)x[5]));
            camera_pose = camera_pose * Tinc;
        }
    }
    kfbSafeCall(ctx, kfb_icp_end(ctx));
    if (!ok) return false;
    return true;
}
