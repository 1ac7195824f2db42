// Header-only stand-in for the handful of OpenCV types the reference's CUDA
// translation units mention.  TEST INFRASTRUCTURE ONLY: it exists so that the
// reference kernels under /root/reference/kfusion/src/*.cu can be compiled
// unmodified for sm_100a into oracle/_ref/ (see oracle/Makefile) and driven by
// oracle/ref_harness.cu as the A/B checker.  Nothing in the product links it.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstring>
#include <cmath>
#include <vector>
#include <string>
#include <iostream>
#include <fstream>

typedef unsigned char uchar;

// Memory-safety interposer for the reference's ICP scratch: rigid_icp.cu:143
// allocates pitch*nblocks BYTES but indexes rows with the pitch as an ELEMENT stride, which runs
// up to ~46 KB past the allocation at pyramid levels 1-2.  The reference source stays unmodified;
// only this allocation is enlarged so the harness cannot fault the GPU.  Results are unaffected.
static inline cudaError_t kfshim_malloc_pitch(void **p, size_t *pitch, size_t width_bytes, size_t height)
{
    const size_t pb = (width_bytes + 511) / 512 * 512;
    *pitch = pb;
    const size_t rows = width_bytes / sizeof(float) + 1; // 27 rows are addressed
    const size_t bytes = (rows * pb + height + 64) * sizeof(float);
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) e = cudaMemset(*p, 0, bytes);
    return e;
}
#define cudaMallocPitch(p, pitch, w, h) kfshim_malloc_pitch((void **)(p), (pitch), (w), (h))

#define CV_8UC3 16
#define CV_32FC1 5
#define CV_32FC3 21

namespace cv
{
    template <typename T, int m, int n>
    struct Matx
    {
        T val[m * n];
        Matx() { for (int i = 0; i < m * n; ++i) val[i] = T(0); }
        T &operator()(int i, int j) { return val[i * n + j]; }
        const T &operator()(int i, int j) const { return val[i * n + j]; }
    };
    template <typename T, int n>
    struct Vec
    {
        T val[n];
        Vec() { for (int i = 0; i < n; ++i) val[i] = T(0); }
        Vec(T a, T b, T c) { val[0] = a; val[1] = b; val[2] = c; }
        T &operator()(int i) { return val[i]; }
        const T &operator()(int i) const { return val[i]; }
        T &operator[](int i) { return val[i]; }
        const T &operator[](int i) const { return val[i]; }
        static Vec all(T v) { Vec r; for (int i = 0; i < n; ++i) r.val[i] = v; return r; }
    };
    typedef Matx<double, 6, 6> Matx66d;
    typedef Matx<float, 3, 3> Matx33f;
    typedef Vec<double, 6> Vec6d;
    typedef Vec<double, 3> Vec3d;
    typedef Vec<float, 3> Vec3f;
    typedef Vec<int, 3> Vec3i;
    typedef Vec<uchar, 3> Vec3b;

    struct Affine3f
    {
        Matx33f R;
        Vec3f t;
        Affine3f() { R(0, 0) = R(1, 1) = R(2, 2) = 1.f; }
        Matx33f rotation() const { return R; }
        Vec3f translation() const { return t; }
    };

    struct Mat
    {
        int rows, cols;
        Mat() : rows(0), cols(0) {}
    };

    namespace cuda
    {
        template <typename T>
        struct PtrSz
        {
            T *data;
            size_t size;
        };
        template <typename T>
        struct PtrStep
        {
            T *data;
            size_t step; // bytes
            __host__ __device__ PtrStep() : data(0), step(0) {}
            __host__ __device__ PtrStep(T *d, size_t s) : data(d), step(s) {}
            __host__ __device__ T *ptr(int y = 0) { return (T *)((char *)data + y * step); }
            __host__ __device__ const T *ptr(int y = 0) const { return (const T *)((const char *)data + y * step); }
            __host__ __device__ T &operator()(int y, int x) { return ptr(y)[x]; }
            __host__ __device__ const T &operator()(int y, int x) const { return ptr(y)[x]; }
        };
        template <typename T>
        struct PtrStepSz : public PtrStep<T>
        {
            int cols, rows;
            __host__ __device__ PtrStepSz() : cols(0), rows(0) {}
            __host__ __device__ PtrStepSz(int r, int c, T *d, size_t s) : PtrStep<T>(d, s), cols(c), rows(r) {}
        };
        // Non-owning view over caller-provided device memory (the harness owns it).
        struct GpuMat
        {
            int rows, cols;
            size_t step;
            unsigned char *data;
            GpuMat() : rows(0), cols(0), step(0), data(0) {}
            GpuMat(int r, int c, size_t elem, void *d) : rows(r), cols(c), step(c * elem), data((unsigned char *)d) {}
            template <typename T> operator PtrStepSz<T>() const { return PtrStepSz<T>(rows, cols, (T *)data, step); }
            template <typename T> operator PtrStep<T>() const { return PtrStep<T>((T *)data, step); }
            void setTo(int) { if (data) cudaMemset(data, 0, step * rows); }
            void release() {}
        };
        inline GpuMat createContinuous(int, int, int) { return GpuMat(); }
    }
}
