/*
 * kf_oracle.h -- CPU restatement of the reference KinectFusion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product (slam-kinectfusion_b200/ and include/).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may use it, and only as the checker / reported baseline.
 *
 * Parity status: the reference (baiyuntao00/SLAM-KinectFusion) ships no
 * tests, golden vectors or fixtures => "parity unpinned"
 * by the reference's own tests.  The restatement is instead pinned against
 * the reference's own CUDA kernels compiled unmodified for sm_100a
 * (oracle/_ref, built by oracle/Makefile) on a B200 (tests/test_ref_ab.py),
 * for every stage except cv::cuda::pyrDown / cv::cuda::bilateralFilter, which
 * are un-vendored OpenCV (opencv_contrib cudawarping / cudaimgproc, version
 * unpinned by the reference, ~4.5.5 by date) and are restated from the
 * published upstream algorithm and cross-checked loosely
 * against cv2 (CPU) in tests/test_oracle.py.
 *
 * Arithmetic convention: every expression is spelled with explicit fmaf()
 * exactly where nvcc 12.9 contracted the reference source for sm_100a
 * (verified in the SASS of oracle/_ref; pattern: in `a*b + c*d [+ e*f]` the
 * left product is fused onto the rounded right product, later products are
 * fused onto the accumulator).  Compile with -ffp-contract=off.  MUFU.RCP
 * (behind __fdividef) is not reproducible on a CPU; KFO_RCP() uses the
 * correctly rounded reciprocal, which is why GPU-vs-oracle tolerances are
 * "1 LSB / few ulp with outlier counts" while GPU-vs-_ref is bit-exact.
 *
 * Conventions shared with include/kfb200.h:
 *   pose12  : float[12] = top three rows of the 4x4, row-major: r00 r01 r02 tx  r10 ... tz
 *   map3    : float3 AoS, 12 B per pixel, row-major (the reference's CV_32FC3 continuous GpuMat)
 *   volume  : int16 pairs {tsdf, weight}, linear index x + y*X + z*X*Y (reference order,
 *             device_utils.cuh:30-37), colour dropped
 */
#ifndef KF_ORACLE_H
#define KF_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    int width, height;
    float fx, fy, cx, cy;
} kfo_intr;

typedef struct
{
    int dims[3];
    float range[3];      /* metres */
    float voxel_size[3]; /* range/dims, computed in float like tsdf_volume.cpp:16 */
    float trunc_dist;
    int max_weight;      /* reference hard-codes 64 (device_utils.cuh:5) */
} kfo_volume_desc;

/* ---- front end ------------------------------------------------------- */
void kfo_level_intrinsics(const kfo_intr *in, int level, kfo_intr *out);
void kfo_pyrdown(const float *src, int w, int h, float *dst, int dw, int dh);
void kfo_bilateral(const float *src, int w, int h, float *dst, int ksize, float sigma_color, float sigma_space);
void kfo_truncate(float *d, int w, int h, float max_dist);
void kfo_vertex_map(const float *d, const kfo_intr *k, float *vmap3);
void kfo_normal_map(const float *vmap3, int w, int h, float *nmap3);
void kfo_resize_maps(const float *vbig, const float *nbig, int bw, int bh, float *vsmall, float *nsmall);

/* ---- ICP --------------------------------------------------------------- */
/* out27: upper triangle of the 6x7 system in reference order (rigid_icp.cu:102-112). */
void kfo_icp_accumulate(const float *cur_v, const float *cur_n, const float *pre_v, const float *pre_n,
                        const kfo_intr *k, const float pose12[12], float dist_thres, float sine_thres,
                        int compat_rows, double out27[27], int64_t *n_corresp);
/* returns 0 on success, 1 on the reference's singular/NaN determinant guard. x6 = [w, t]. */
int kfo_icp_solve(const double in27[27], double x6[6]);
void kfo_pose_apply_increment(float pose12[12], const double x6[6]); /* pose = pose * Tinc(x) */

/* ---- pose algebra (restated OpenCV core) ------------------ */
void kfo_pose_identity(float p[12]);
void kfo_pose_mul(const float a[12], const float b[12], float out[12]);
void kfo_pose_inv(const float a[12], float out[12]);
void kfo_rot_inv(const float a[12], float rinv9[9]);
void kfo_rodrigues(const float rvec[3], float R9[9]);

/* ---- TSDF ------------------------------------------------------------- */
void kfo_integrate(int16_t *vol, const kfo_volume_desc *vd, const float vol2cam[12],
                   const float *depth_m, const kfo_intr *k, int z_begin, int z_end, int64_t *n_updated);
void kfo_raycast(const int16_t *vol, const kfo_volume_desc *vd, const float cam2vol[12], const float rinv9[9],
                 const kfo_intr *k, float *vmap3, float *nmap3, int64_t *n_steps, int compat_ts_sign);
/* z-slab variant (test model of the sharded raycast): see kf_oracle.c */
void kfo_raycast_slab(const int16_t *vol, const kfo_volume_desc *vd, const float cam2vol[12], const float rinv9[9],
                      const kfo_intr *k, float *vmap3, float *nmap3, float *key, int zs0, int zs1, int zo0, int zo1,
                      int compat_ts_sign);
/* Point-cloud extraction (tsdf_volume.cu:307-481); returns the number of points written (<= cap). */
int64_t kfo_extract_points(const int16_t *vol, const kfo_volume_desc *vd, const float volpose[12],
                           float *points3, int64_t cap);
void kfo_render_phong(const float *vmap3, const float *nmap3, int w, int h, const float eye[3], uint8_t *bgr);
void kfo_render_normals(const float *nmap3, int w, int h, uint8_t *bgr);

/* ---- synthetic scene ---------------------------------------- */
void kfo_trajectory_pose(int k, int period, float pose12[12]);
void kfo_render_depth_mm(const float cam2world[12], const kfo_intr *k, float *depth_mm);
/* wide-FOV wall frame for the dense-update micro-config (constant z-depth, mm). */
void kfo_fill_const_depth_mm(int w, int h, float mm, float *depth_mm);

/* ---- whole pipeline (kinectfusion.cpp:48-127 state machine) --------------- */
typedef struct kfo_kinfu kfo_kinfu;
typedef struct
{
    int pyramid_height;          /* 3 */
    float dfilter_dist;          /* 5.0 */
    int bfilter_kernel_size;     /* 5 */
    float bfilter_spatial_sigma; /* 10 */
    float bfilter_color_sigma;   /* 10 */
    float icp_dist_threshold;    /* 0.015 */
    float icp_angle_threshold;   /* degrees, 30 */
    int icp_iter_count[8];       /* indexed by level: {4,5,10} */
    float volu_range[3];
    float volu_pose[12];
    float volu_trun_dist;
    int volu_dims[3];
    int tsdf_max_weight;
    int compat_icp_rows;         /* 1 = reference's truncated grid */
    int compat_raycast_ts_sign;  /* 1 = reference's minus sign */
} kfo_params;
void kfo_default_params(kfo_params *p, int dims);
kfo_kinfu *kfo_kinfu_create(const kfo_intr *k, const kfo_params *p);
void kfo_kinfu_destroy(kfo_kinfu *kf);
void kfo_kinfu_reset(kfo_kinfu *kf);
/* returns 0 ok, 1 tracking failure (=> reset done, like kinectfusion.cpp:97-102) */
int kfo_kinfu_pipeline(kfo_kinfu *kf, const float *depth_mm);
int kfo_kinfu_frame_count(const kfo_kinfu *kf);
int kfo_kinfu_num_poses(const kfo_kinfu *kf);
void kfo_kinfu_get_pose(const kfo_kinfu *kf, int idx, float pose12[12]); /* idx<0 => last */
int16_t *kfo_kinfu_volume(kfo_kinfu *kf);
const float *kfo_kinfu_cur_depth(const kfo_kinfu *kf, int level);
const float *kfo_kinfu_cur_vmap(const kfo_kinfu *kf, int level);
const float *kfo_kinfu_cur_nmap(const kfo_kinfu *kf, int level);
const float *kfo_kinfu_prev_vmap(const kfo_kinfu *kf, int level);
const float *kfo_kinfu_prev_nmap(const kfo_kinfu *kf, int level);
int64_t kfo_kinfu_last_updated(const kfo_kinfu *kf);
int64_t kfo_kinfu_last_raysteps(const kfo_kinfu *kf);
/* per-stage wall seconds of the last pipeline() call: frontend, icp, integrate, raycast */
void kfo_kinfu_last_times(const kfo_kinfu *kf, double t4[4]);

#ifdef __cplusplus
}
#endif
#endif
