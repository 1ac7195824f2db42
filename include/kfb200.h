/*
 * kfb200.h -- C-ABI of the B200-native KinectFusion tracking-and-mapping core.
 *
 * This is the drop-in boundary.  It replaces, one for one, the
 * host->device seam of baiyuntao00/SLAM-KinectFusion: the eleven free functions
 * `kf::device::*` declared at kfusion/include/device_types.hpp:113-128 plus the
 * two OpenCV-CUDA calls and the GpuMat upload/download/setTo traffic in
 * kfusion/src/kinectfusion.cpp:39-65.  Each entry point cites the reference
 * interface it replaces.  Plain C: opaque context, plain pointers and sizes,
 * int return codes (0 = ok), no C++/torch/OpenCV types.
 *
 * Conventions
 *   pose12 : float[12], the top three rows of the 4x4 matrix, row-major
 *            (r00 r01 r02 tx r10 r11 r12 ty r20 r21 r22 tz) == cv::Affine3f::matrix.val[0..11]
 *   map3   : float3 array-of-structs, 12 B per pixel, row-major (the reference's
 *            continuous CV_32FC3 GpuMat, types.hpp:45-46); device storage is float4
 *   volume : packed {int16 tsdf, int16 weight} voxels; the reference's 3 colour
 *            bytes are write-only dead state and are dropped.
 *            Hosts exchange volumes in the reference's linear order x + y*X + z*X*Y
 *            (device_utils.cuh:30-37); in HBM the voxels live in 8x8x8 bricks of 2 KB
 *            (bricks in x, y, z order, voxels inside a brick in x, y, z order, dims
 *            padded to whole bricks): the layout the raycaster's gathers and the
 *            sweep's patches want.  dims[0] % 4 == 0 is required (128-bit accesses).
 *   Threading: one host thread per context; all work runs on a context-owned
 *            non-blocking CUDA stream.  Functions that return host data block
 *            until that data is valid; the others only enqueue.
 *   There is NO CPU fallback: every compute entry point launches sm_100a kernels
 *            and fails with KFB_ERR_CUDA if no such device is present.
 */
#ifndef KFB200_H
#define KFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KFB_MAX_LEVELS 8

enum
{
    KFB_OK = 0,
    KFB_ERR_INVALID = 1,     /* bad argument */
    KFB_ERR_CUDA = 2,        /* CUDA runtime error, see kfb_last_error_string */
    KFB_ERR_UNSUPPORTED = 3, /* e.g. dims[0] % 4 != 0 */
    KFB_ERR_TRACKING = 4,    /* reserved for host facades (icp_registration.cpp:35-37) */
    KFB_ERR_TIMEOUT = 5      /* a host<->device handshake timed out (not a tracking failure; the map is intact) */
};

/* kf::Intrinsics, kfusion/include/types.hpp:13-29 (the unused depth scale `c` omitted). */
typedef struct
{
    int width, height;
    float fx, fy, cx, cy;
} kfb_intrinsics;

/* kf::kinectfuison_params, kfusion/include/kinectfusion.h:9-30; defaults from
 * kfusion/src/kinectfusion.cpp:167-190 via kfb_default_params(). */
typedef struct
{
    int pyramid_height;            /* 3 */
    float dfilter_dist;            /* 5.0 m */
    int bfilter_kernel_size;       /* 5 */
    float bfilter_spatial_sigma;   /* 10 */
    float bfilter_color_sigma;     /* 10 (mm) */
    float icp_dist_threshold;      /* 0.015 m */
    float icp_angle_threshold;     /* 30 degrees */
    int icp_iter_count[KFB_MAX_LEVELS]; /* indexed by level: {4,5,10} */
    int volu_dims[3];              /* 512^3 */
    float volu_range[3];           /* 3 m */
    float volu_trun_dist;          /* 2.1 * range/dims */
    int tsdf_max_weight;           /* 64 (the reference hard-codes MAX_WEIGHT, device_utils.cuh:5) */
    int compat_icp_rows;           /* 1: reproduce the reference's truncated ICP grid (rigid_icp.cu:137-139) */
    int compat_raycast_ts_sign;    /* 1: reproduce the reference's Ts sign (tsdf_volume.cu:246) */
    /* z-slab sharding for volumes of 1024^3 and above: this context
     * stores and updates planes [slab_z_begin - halo, slab_z_end + halo) only.
     * 0,0 = the whole volume. */
    int slab_z_begin, slab_z_end;
} kfb_params;

typedef struct kfb_ctx kfb_ctx;

/* ---- lifecycle ---------------------------------------------------------- */
void kfb_default_params(kfb_params *p, int dims);                       /* kinectfusion.cpp:167-190 */
int kfb_create(const kfb_intrinsics *intr, const kfb_params *p, int device, kfb_ctx **out);
                                                                        /* kinectfusion.cpp:9-27, types.hpp:37-52, tsdf_volume.cpp:13-27 */
void kfb_destroy(kfb_ctx *ctx);                                         /* kinectfusion.cpp:191-196 */
const char *kfb_last_error_string(const kfb_ctx *ctx);                  /* safe_call.hpp:8-14 (returned, not printed) */
int kfb_synchronize(kfb_ctx *ctx);                                      /* cudaDeviceSynchronize, tsdf_volume.cu:110 */
int kfb_device_count(void);
/* Run all further work of this context on a caller-owned CUDA stream (cudaStream_t), e.g. the stream a
 * torch.distributed / NCCL communicator orders its collectives against.  The reference runs everything
 * on the legacy default stream (all `<<<grid, block>>>` launches); this is the explicit equivalent. */
int kfb_set_stream(kfb_ctx *ctx, void *cuda_stream);

/* ---- volume ------------------------------------------------------------- */
int kfb_reset_volume(kfb_ctx *ctx);                                     /* device::resetVolume, device_types.hpp:116 */
int kfb_reset_frames(kfb_ctx *ctx);                                     /* Frame::reset, types.hpp:53-62 */

/* ---- frame ingest + front end ---------------------------------------------- */
/* GpuMat::upload of the f32 millimetre depth, kinectfusion.cpp:50.  `host` may be pageable or
 * pinned host memory (asynchronous on the context stream when pinned) or a device pointer
 * (frame already resident in HBM: device-to-device copy). */
int kfb_upload_depth_mm(kfb_ctx *ctx, const float *host, int width, int height);
/* The same from the sensor's native 16-bit millimetre image (what depth_sensor.cpp:186-196 converts to float on
 * the host before the upload): half the PCIe bytes, the conversion to f32 runs on the device.  Pinned host or
 * device pointers are used in place, pageable ones are staged. */
int kfb_upload_depth_mm_u16(kfb_ctx *ctx, const uint16_t *host, int width, int height);
/* Buffer lifetime: from pinned host memory (or a device pointer) the upload is asynchronous and reads the caller's buffer
 * after the call has returned; kfb_upload_wait blocks until the last upload has left that buffer, i.e. until it may be
 * reused (GpuMat::upload is synchronous, kinectfusion.cpp:50).  kf::kinectfusion::pipeline calls it before it returns
 * on the frames that did not wait for the device anyway (the bootstrap frame; slab ranks other than 0). */
int kfb_upload_wait(kfb_ctx *ctx);
/* cv::cuda::pyrDown x(L-1), cv::cuda::bilateralFilter xL, device::depthTruncation xL,
 * device::getVertexmap xL, device::getNormalmap xL -- kinectfusion.cpp:54-75,
 * device_types.hpp:122-124.  Fills the CURRENT frame's depth/vertex/normal pyramids. */
int kfb_frontend(kfb_ctx *ctx);
/* cframe->vmap.swap(pframe->vmap); cframe->nmap.swap(pframe->nmap) -- kinectfusion.cpp:88-89 */
int kfb_swap_frames(kfb_ctx *ctx);

/* ---- ICP ------------------------------------------------------------------ */
/* device::rigidICP (device_types.hpp:127, rigid_icp.cu:135-169) for one pyramid level:
 * projective association of the current maps (under pose12 = current estimate of
 * cur->prev) against the previous/model maps, and the 27 unique sums of the 6x7
 * normal equations in the reference's order.  Blocks until out27 is valid. */
int kfb_icp_accumulate(kfb_ctx *ctx, int level, const float pose12[12], double out27[27]);
/* The same operation for the whole coarse-to-fine loop of ICPRegistration::rigidTransform
 * (icp_registration.cpp:21-43) without per-iteration launch latency: kfb_icp_begin declares the
 * schedule (iters_per_level[l] iterations at level l, coarsest level first; at most 255 in all); the first
 * kfb_icp_step starts ONE kernel that runs every iteration by itself: it computes each next pose with the
 * reference's update rule (icp_registration.cpp:35-42: solve, Rodrigues, pose * Tinc, in the facade's arithmetic)
 * and posts every iteration's 27 sums together with the pose it used into mapped host memory.  kfb_icp_step(k)
 * waits for iteration k's slot, compares that pose with pose12 bit for bit and returns the sums -- exactly what
 * kfb_icp_accumulate would return for pose12.  A caller whose update rule gives another pose is still served
 * correctly: from the first difference on, the schedule (and the context's next 64) runs as one ordinary launch
 * per step (kfb_icp_mispredict_count / kfb_icp_fallback_count tell).  KFB_ICP_DIRECT=1 forces that mode.
 * kfb_icp_end closes the schedule; iterations never stepped (tracking failure) cost nothing on the host. */
int kfb_icp_begin(kfb_ctx *ctx, const int iters_per_level[KFB_MAX_LEVELS]);
int kfb_icp_step(kfb_ctx *ctx, const float pose12[12], double out27[27]);
int kfb_icp_end(kfb_ctx *ctx);

/* ---- TSDF ---------------------------------------------------------------------- */
/* device::integrate (device_types.hpp:118, tsdf_volume.cu:103-111) with the current
 * frame's level-0 filtered depth.  vol2cam12 = camera_pose.inv() * volume_pose
 * (tsdf_volume.cpp:50).  n_updated may be NULL; when non-NULL a counting variant of
 * the kernel runs and the call blocks until the count is valid. */
int kfb_integrate(kfb_ctx *ctx, const float vol2cam12[12], uint64_t *n_updated);
/* Slab balancing aid for sharded volumes (no reference counterpart): host_hist[z], z in [0, dims[2]), = work the
 * sweep of kfb_integrate would do on plane z of the WHOLE volume for the current frame's depth and this pose, in
 * voxel quads: every plane a work item covers counts its 32 quads once if the item only streams free space and ten
 * times if it needs the per-voxel predicate (the sweep's measured cost ratio), plus the raycast's share: 190 quads'
 * worth per pixel with a valid depth, spread over the 16 planes in front of the pixel's surface point.  It comes from the sweep's own plan, which
 * depends on the depth image and the pose only -- not on the volume's content or on the planes this context stores
 * -- so every rank computes the same histogram without communication. */
int kfb_integrate_plane_histogram(kfb_ctx *ctx, const float vol2cam12[12], uint32_t *host_hist);
/* device::raycast (device_types.hpp:117, tsdf_volume.cu:264-273) into the PREVIOUS
 * (model) frame's level-0 maps, misses written as zeros (pframe->reset(),
 * kinectfusion.cpp:112).  cam2vol12 = volume_pose.inv()*camera_pose and
 * rinv9 = its rotation inverted, row-major (tsdf_volume.cpp:59-61). */
int kfb_raycast(kfb_ctx *ctx, const float cam2vol12[12], const float rinv9[9]);
/* device::resizePointsNormals x(L-1) on the model maps, kinectfusion.cpp:114-120. */
int kfb_model_pyramid(kfb_ctx *ctx);
/* z-slab sharded volumes (no reference counterpart, the reference is single-GPU): a slab
 * context's kfb_raycast marches only the ray samples whose voxel lies in its slab and records, per pixel,
 * the ray length of its first terminal event (hit or back-face stop; +inf if none) in the key buffer
 * kfb_device_ptr(ctx, 4).  After the per-pixel minimum of the keys over all slabs has been formed
 * (all-reduce MIN), kfb_composite_mask zeroes this context's model vertex/normal maps wherever it does not
 * hold the winning key, so that an integer SUM reduction of the maps over the slabs yields exactly the
 * single-GPU raycast. */
int kfb_composite_mask(kfb_ctx *ctx, const float *min_key_dev);
/* The same composite without a communication library, over NVLink peer memory (one process per GPU on one NVLink
 * box): every slab context exports CUDA IPC handles of the buffers its peers may write into -- staging for event
 * keys, staging for vertex + normal maps (one slot per rank) and a set of "slab done" counters (kfb_ipc_export,
 * `which` = 0..3, 64 bytes each; 3 repeats 2); the launcher gathers them over any channel and hands every context the
 * full table (kfb_shard_attach: world x 4 handles, rank-major).  From then on the raycast of an attached rank other
 * than 0 PUSHES its keys and maps into its slot of rank 0's staging (posted NVLink stores from the kernel's epilogue),
 * and kfb_shard_composite, called by every rank after kfb_raycast, raises the rank's counter in rank 0's memory and,
 * on rank 0, runs ONE kernel that waits for all counters, selects the first terminal event per pixel from local
 * memory, writes the winner's vertex and normal into rank 0's model maps (bit-identical to the single-GPU raycast)
 * and derives the coarser pyramid levels from the same tiles.  A peer that never signals within 2 s is reported
 * (KFB_ERR_TIMEOUT from the next blocking call) instead of compositing stale data.  Slot r of the staging is written
 * by rank r only, once per frame, after rank 0 consumed the previous frame's (implied by waiting for the next pose). */
#define KFB_IPC_HANDLE_BYTES 64
int kfb_ipc_export(kfb_ctx *ctx, int which, void *handle64);
int kfb_shard_attach(kfb_ctx *ctx, int rank, int world, const void *handles);
int kfb_shard_composite(kfb_ctx *ctx);
int kfb_shard_attached(const kfb_ctx *ctx);
/* Unmap the peers' buffers again (waits for this context's stream first).  Every rank must have detached before any
 * rank destroys its context: freeing memory a peer still has mapped is undefined. */
int kfb_shard_detach(kfb_ctx *ctx);

/* ---- export ("next" rows) ------------------------------------------ */
/* device::extract_points (device_types.hpp:128, tsdf_volume.cu:483-499) + the D2H of
 * TSDFVolume::fetchPointCloud (tsdf_volume.cpp:63-84).  Writes up to `cap` xyz points
 * (world frame via volpose12) to host_points3; *n_points = number written. */
int kfb_extract_points(kfb_ctx *ctx, const float volpose12[12], float *host_points3, size_t cap, size_t *n_points);
/* device::renderPhong / device::renderNormals (device_types.hpp:120-121) of the model
 * maps + GpuMat::download (kinectfusion.cpp:33-47).  host_bgr: width*height*3 bytes. */
int kfb_render_phong(kfb_ctx *ctx, const float eye3[3], uint8_t *host_bgr);
int kfb_render_normals(kfb_ctx *ctx, uint8_t *host_bgr);

/* ---- test / interop hooks (GpuMat::download/upload equivalents) --------------------- */
enum { KFB_FRAME_CUR = 0, KFB_FRAME_PREV = 1 };
int kfb_download_depth(kfb_ctx *ctx, int level, float *host);                 /* current frame, metres */
int kfb_upload_depth_m(kfb_ctx *ctx, int level, const float *host);           /* inject filtered depth */
int kfb_download_raw_depth(kfb_ctx *ctx, int level, float *host);             /* pyrDown chain, millimetres */
int kfb_download_maps(kfb_ctx *ctx, int frame, int level, float *host_v3, float *host_n3);
int kfb_upload_maps(kfb_ctx *ctx, int frame, int level, const float *host_v3, const float *host_n3);
/* whole (slab of the) volume as int16 pairs in reference index order */
int kfb_download_volume(kfb_ctx *ctx, int16_t *host);
int kfb_upload_volume(kfb_ctx *ctx, const int16_t *host);
size_t kfb_volume_voxels(const kfb_ctx *ctx);                                  /* voxels stored by this context */
void kfb_level_intrinsics(const kfb_intrinsics *in, int level, kfb_intrinsics *out); /* types.hpp:18-28 */

/* ---- measurement ------------------------------------------------------------------- */
/* cudaEvent pool on the context stream: record slot i now; elapsed ms between slots. */
int kfb_event_record(kfb_ctx *ctx, int slot);
int kfb_event_elapsed_ms(kfb_ctx *ctx, int slot_a, int slot_b, float *ms);
/* opt-in stage profiling: when on, launchers bracket their main kernel with events in reserved
 * slots (whole-schedule ICP kernel: 54/55, shard composite kernel: 52/53, integrate kernel: 60/61, whole kfb_integrate call: 56/57, raycast: 58/59) so a caller can read that
 * kernel's own duration. */
int kfb_set_profiling(kfb_ctx *ctx, int on);
/* number of kernels this library has launched on this context since creation */
uint64_t kfb_launch_count(const kfb_ctx *ctx);
/* raw device pointers for zero-copy interop (NCCL / torch views); which: 0 volume (brick-major, see "volume" above),
 * 1 prev vmap L0 (float4; the nmap follows it contiguously), 2 prev nmap L0 (float4), 3 cur depth L0 (float),
 * 4 raycast event keys (float) */
void *kfb_device_ptr(kfb_ctx *ctx, int which);
void *kfb_stream(kfb_ctx *ctx);
/* debug: %globaltimer (ns) per iteration of the last whole-schedule ICP kernel, as seen by CTA 0 (row = iteration
 * index % 32): {iteration entry, pixels accumulated, final sums ready, next pose ready} */
void kfb_debug_icp_ring(kfb_ctx *ctx, uint64_t out128[128]);
/* number of ICP schedules (or rests of schedules) that ran as ordinary per-iteration launches because the
 * whole-schedule kernel could not be made co-resident, gave up on a poll, or had predicted another pose than the
 * caller's (results are bit-identical either way) */
uint64_t kfb_icp_fallback_count(const kfb_ctx *ctx);
/* number of kfb_icp_step calls whose pose differed from the one the free-running kernel had predicted for that
 * iteration (each one ends the free run of its schedule; 0 for callers that use the reference's update rule) */
uint64_t kfb_icp_mispredict_count(const kfb_ctx *ctx);
/* debug / measurement: numbers of the last COUNTING kfb_integrate call (n_updated != NULL): {updated voxels, 16-byte
 * voxel quads loaded, quads stored, stream work items, general work items, 0}.  Loads + stores x 16 B is the
 * kernel's own count of the bytes it moved (stores of unchanged values are dropped). */
void kfb_debug_integrate_counts(kfb_ctx *ctx, uint64_t out6[6]);

#ifdef __cplusplus
}
#endif
#endif /* KFB200_H */
