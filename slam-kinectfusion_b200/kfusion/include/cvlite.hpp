// cvlite.hpp -- the handful of OpenCV value types the reference's public surface mentions
// (cv::Vec3f, cv::Vec3i, cv::Matx33f, cv::Affine3f, cv::Mat), re-expressed as small PODs so the
// host facade builds without OpenCV.  Semantics restate OpenCV core:
// Affine3f(rvec, t) = Rodrigues in double on float rvec; inv() = general affine inverse evaluated
// in double; operator* = 4x4 float product.  Define KF_NO_CV_ALIAS to keep the `cv` namespace free
// (e.g. when real OpenCV headers are also included).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace kfcv
{
template <typename T, int n>
struct Vec
{
    T val[n];
    Vec() { for (int i = 0; i < n; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(n == 3, "3-vector"); val[0] = a; val[1] = b; val[2] = c; }
    T &operator()(int i) { return val[i]; }
    const T &operator()(int i) const { return val[i]; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
    static Vec all(T v) { Vec r; for (int i = 0; i < n; ++i) r.val[i] = v; return r; }
};
typedef Vec<float, 3> Vec3f;
typedef Vec<int, 3> Vec3i;
typedef Vec<double, 3> Vec3d;
typedef Vec<double, 6> Vec6d;
typedef Vec<uint8_t, 3> Vec3b;

template <typename T, int m, int n>
struct Matx
{
    T val[m * n];
    Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
    T &operator()(int i, int j) { return val[i * n + j]; }
    const T &operator()(int i, int j) const { return val[i * n + j]; }
};
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<float, 4, 4> Matx44f;
typedef Matx<double, 6, 6> Matx66d;

struct Affine3f
{
    Matx44f matrix;
    Affine3f() { matrix(0, 0) = matrix(1, 1) = matrix(2, 2) = matrix(3, 3) = 1.f; }
    Affine3f(const Matx33f &R, const Vec3f &t) : Affine3f()
    {
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) matrix(i, j) = R(i, j); matrix(i, 3) = t(i); }
    }
    // cv::Affine3f(rvec, t): Rodrigues
    Affine3f(const Vec3f &rvec, const Vec3f &t) : Affine3f()
    {
        const double rx = rvec(0), ry = rvec(1), rz = rvec(2);
        const double theta = std::sqrt(rx * rx + ry * ry + rz * rz);
        if (theta >= DBL_EPSILON)
        {
            const double c = std::cos(theta), s = std::sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
            const double x = rx * it, y = ry * it, z = rz * it;
            matrix(0, 0) = (float)(c + c1 * x * x);     matrix(0, 1) = (float)(c1 * x * y - s * z); matrix(0, 2) = (float)(c1 * x * z + s * y);
            matrix(1, 0) = (float)(c1 * x * y + s * z); matrix(1, 1) = (float)(c + c1 * y * y);     matrix(1, 2) = (float)(c1 * y * z - s * x);
            matrix(2, 0) = (float)(c1 * x * z - s * y); matrix(2, 1) = (float)(c1 * y * z + s * x); matrix(2, 2) = (float)(c + c1 * z * z);
        }
        for (int i = 0; i < 3; ++i) matrix(i, 3) = t(i);
    }
    static Affine3f Identity() { return Affine3f(); }
    Matx33f rotation() const { Matx33f R; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R(i, j) = matrix(i, j); return R; }
    Vec3f translation() const { return Vec3f(matrix(0, 3), matrix(1, 3), matrix(2, 3)); }
    Affine3f translate(const Vec3f &t) const
    {
        Affine3f r = *this;
        for (int i = 0; i < 3; ++i) r.matrix(i, 3) += t(i);
        return r;
    }
    Affine3f inv() const
    {
        const double m00 = matrix(0, 0), m01 = matrix(0, 1), m02 = matrix(0, 2), m10 = matrix(1, 0), m11 = matrix(1, 1),
                     m12 = matrix(1, 2), m20 = matrix(2, 0), m21 = matrix(2, 1), m22 = matrix(2, 2);
        const double c00 = m11 * m22 - m12 * m21, c01 = m12 * m20 - m10 * m22, c02 = m10 * m21 - m11 * m20;
        const double det = m00 * c00 + m01 * c01 + m02 * c02, id = 1.0 / det;
        double iv[9];
        iv[0] = c00 * id; iv[1] = (m02 * m21 - m01 * m22) * id; iv[2] = (m01 * m12 - m02 * m11) * id;
        iv[3] = c01 * id; iv[4] = (m00 * m22 - m02 * m20) * id; iv[5] = (m02 * m10 - m00 * m12) * id;
        iv[6] = c02 * id; iv[7] = (m01 * m20 - m00 * m21) * id; iv[8] = (m00 * m11 - m01 * m10) * id;
        const double tx = matrix(0, 3), ty = matrix(1, 3), tz = matrix(2, 3);
        Affine3f r;
        for (int i = 0; i < 3; ++i)
        {
            for (int j = 0; j < 3; ++j) r.matrix(i, j) = (float)iv[3 * i + j];
            r.matrix(i, 3) = (float)(-(iv[3 * i] * tx + iv[3 * i + 1] * ty + iv[3 * i + 2] * tz));
        }
        return r;
    }
    // first three rows, row-major: the pose12 of include/kfb200.h
    void to12(float p[12]) const { std::memcpy(p, matrix.val, 12 * sizeof(float)); }
    static Affine3f from12(const float p[12]) { Affine3f a; std::memcpy(a.matrix.val, p, 12 * sizeof(float)); return a; }
};
inline Affine3f operator*(const Affine3f &a, const Affine3f &b)
{
    Affine3f r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j)
        {
            float s = 0.f;
            for (int q = 0; q < 3; ++q) s += a.matrix(i, q) * b.matrix(q, j);
            if (j == 3) s += a.matrix(i, 3);
            r.matrix(i, j) = s;
        }
    return r;
}

enum { CV_8UC3 = 16, CV_32FC1 = 5, CV_32FC3 = 21 };
// Minimal owning host image (row-major, continuous) standing in for cv::Mat / cv::Mat_<float>.
struct Mat
{
    int rows, cols, type_;
    std::vector<uint8_t> store;
    Mat() : rows(0), cols(0), type_(CV_32FC1) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), store((size_t)r * c * elemSize1(type)) {}
    Mat(int r, int c, int type, const void *src) : Mat(r, c, type) { std::memcpy(store.data(), src, store.size()); }
    static size_t elemSize1(int type) { return type == CV_8UC3 ? 3 : (type == CV_32FC3 ? 12 : 4); }
    size_t elemSize() const { return elemSize1(type_); }
    int type() const { return type_; }
    bool empty() const { return store.empty(); }
    template <typename T> T *ptr(int y = 0) { return reinterpret_cast<T *>(store.data() + (size_t)y * cols * elemSize()); }
    template <typename T> const T *ptr(int y = 0) const { return reinterpret_cast<const T *>(store.data() + (size_t)y * cols * elemSize()); }
    template <typename T> T &at(int y, int x) { return ptr<T>(y)[x]; }
    template <typename T> const T &at(int y, int x) const { return ptr<T>(y)[x]; }
    void setTo(int v) { std::memset(store.data(), v, store.size()); }
};
} // namespace kfcv

#ifndef KF_NO_CV_ALIAS
namespace cv = kfcv;
#endif
