// kfb_api.cu -- the C-ABI of include/kfb200.h: context lifetime, frame ingest, test hooks,
// measurement helpers.  Compute entry points forward to the per-stage launchers.
#include "kfb_common.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

using namespace kfb;

static thread_local std::string g_create_err;

namespace kfb
{
int join_front(kfb_ctx *ctx)
{
    if (ctx->front_pending)
    {
        KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_front, 0));
        ctx->front_pending = 0;
    }
    return KFB_OK;
}
int fork_front(kfb_ctx *ctx)
{
    KFB_CUDA(ctx, cudaStreamWaitEvent(ctx->fstream, ctx->ev_free, 0));
    return KFB_OK;
}
int mark_free(kfb_ctx *ctx)
{
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_free, ctx->stream));
    return KFB_OK;
}
} // namespace kfb

extern "C" {

void kfb_level_intrinsics(const kfb_intrinsics *in, int level, kfb_intrinsics *out)
{
    // kf::Intrinsics::level, kfusion/include/types.hpp:18-28
    if (level == 0) { *out = *in; return; }
    const float s = powf(0.5f, (float)level);
    out->width = in->width >> level;
    out->height = in->height >> level;
    out->fx = in->fx * s;
    out->fy = in->fy * s;
    out->cx = (in->cx + 0.5f) * s - 0.5f;
    out->cy = (in->cy + 0.5f) * s - 0.5f;
}

void kfb_default_params(kfb_params *p, int dims)
{
    // kf::kinectfuison_params::default_params, kfusion/src/kinectfusion.cpp:167-190
    memset(p, 0, sizeof(*p));
    p->pyramid_height = 3;
    p->bfilter_color_sigma = 10;
    p->bfilter_spatial_sigma = 10;
    p->bfilter_kernel_size = 5;
    p->dfilter_dist = 5.f;
    p->icp_angle_threshold = 30.f;
    p->icp_dist_threshold = 0.015f;
    p->icp_iter_count[0] = 4; p->icp_iter_count[1] = 5; p->icp_iter_count[2] = 10;
    for (int i = 0; i < 3; ++i) { p->volu_dims[i] = dims; p->volu_range[i] = 3.f; }
    p->volu_trun_dist = 2.1f * p->volu_range[0] / (float)p->volu_dims[0];
    p->tsdf_max_weight = 64;
    p->compat_icp_rows = 1;
    p->compat_raycast_ts_sign = 1;
    p->slab_z_begin = p->slab_z_end = 0;
}

int kfb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

#define KFB_HALO 3 // planes integrated redundantly on each side of a slab

int kfb_create(const kfb_intrinsics *intr, const kfb_params *p, int device, kfb_ctx **out)
{
    if (!intr || !p || !out) return KFB_ERR_INVALID;
    *out = nullptr;
    if (p->pyramid_height < 1 || p->pyramid_height > KFB_MAX_LEVELS) return KFB_ERR_INVALID;
    if (intr->width <= 0 || intr->height <= 0) return KFB_ERR_INVALID;
    for (int i = 0; i < 3; ++i) if (p->volu_dims[i] < 4) return KFB_ERR_INVALID;
    if (p->volu_dims[0] % 4 != 0) return KFB_ERR_UNSUPPORTED; // 128-bit voxel-row accesses
    if ((intr->width >> (p->pyramid_height - 1)) < 3 || (intr->height >> (p->pyramid_height - 1)) < 3) return KFB_ERR_INVALID;
    kfb_ctx *ctx = new (std::nothrow) kfb_ctx();
    if (!ctx) return KFB_ERR_INVALID;
    ctx->device = device;
    ctx->intr = *intr;
    ctx->p = *p;
    ctx->levels = p->pyramid_height;
    ctx->cur = 0; ctx->prev = 1;
    ctx->launches = 0;
    ctx->profiling = 0;
    ctx->icp_seq = 0;
    ctx->pyramid_fresh = 0;
    ctx->shard_rank = 0; ctx->shard_world = 0; ctx->shard_flag = nullptr; ctx->shard_seq = 0; ctx->stage_keys = nullptr; ctx->stage_maps = nullptr;
    memset(ctx->peer_keys, 0, sizeof(ctx->peer_keys)); memset(ctx->peer_maps, 0, sizeof(ctx->peer_maps)); memset(ctx->peer_flag, 0, sizeof(ctx->peer_flag));
    ctx->vol = nullptr; ctx->cloud = nullptr; ctx->cloud_cap = 0;
    ctx->tab_thrz = nullptr; ctx->wtab = nullptr; ctx->zexit = nullptr; ctx->zsparse = nullptr; ctx->bricks = nullptr;
    ctx->tab4 = nullptr; ctx->plan_buf = nullptr; ctx->plan_bytes = 0; ctx->plan_hint_host = nullptr; ctx->gen_attr_set = 0; ctx->integrate_seq = 0;
    ctx->bdist = ctx->bdist_tmp = ctx->bdist_tmp2 = nullptr; ctx->bdirty = nullptr; ctx->bdirty_tag = 0;
    ctx->ray_cost = nullptr; ctx->ray_order = nullptr; ctx->ray_order_valid = 0; ctx->ev_ray_done = nullptr; ctx->ev_ray_order = nullptr; ctx->ev_order_gate = nullptr;
    ctx->ray_order_pending = 0; ctx->ray_order_tiles = 0; ctx->ostream = nullptr;
    ctx->hit_t = nullptr; ctx->icp_partials = nullptr; ctx->icp_ticket = nullptr; ctx->counters = nullptr;
    ctx->counters_host = nullptr; ctx->pinned_depth = nullptr; ctx->depth_u16 = nullptr; ctx->render_dev = nullptr; ctx->render_host = nullptr;
    ctx->stream = nullptr; ctx->own_stream = 1;
    ctx->fstream = nullptr; ctx->ev_upload = nullptr; ctx->ev_front = nullptr; ctx->ev_free = nullptr; ctx->ev_tables_free = nullptr; ctx->front_pending = 0;
    ctx->icp_host = nullptr; ctx->icp_devgate = nullptr;
    ctx->icp_smem_set = 0;
    ctx->icp_fallbacks = 0; ctx->icp_direct_left = 0;
    ctx->icp_slots_host = ctx->icp_slots_dev = nullptr; ctx->icp_tagged = nullptr; ctx->icp_tagged_cap = 0; ctx->icp_mispredicts = 0;
    ctx->istream = nullptr; ctx->ev_ifork = nullptr; ctx->ev_ijoin = nullptr; ctx->ev_plan_clean = nullptr; ctx->ev_sweep_main = nullptr; ctx->plan_clean_bytes = 0;
    ctx->dev_err_host = ctx->dev_err_dev = nullptr;
    memset(ctx->L, 0, sizeof(ctx->L));
    memset(ctx->events, 0, sizeof(ctx->events));
    *out = ctx; // returned even on failure so the caller can read the error string, then destroy
    KFB_CUDA(ctx, cudaSetDevice(device));
    {
        cudaDeviceProp prop;
        KFB_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) { ctx->err = "kfb200 needs an sm_100a device (compute capability 10.x)"; return KFB_ERR_CUDA; }
        ctx->sm_count = prop.multiProcessorCount;
    }
    {
        int lo = 0, hi = 0;
        KFB_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        // three levels: the general sweep kernel's side stream above the frame's main stream above the next frame's
        // front end (numerically lower = more urgent; measured: the general kernel losing SM slots to the streaming
        // kernel costs 11 % of the sweep, the main stream yielding to the front end 0.5 % of the frame)
        KFB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, (lo + hi) / 2));
        KFB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->fstream, cudaStreamNonBlocking));
        KFB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->istream, cudaStreamNonBlocking, hi));
        KFB_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->ostream, cudaStreamNonBlocking, lo));
    }
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ifork, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_ijoin, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_plan_clean, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_sweep_main, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_upload, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_upload, ctx->fstream));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_front, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_free, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_tables_free, cudaEventDisableTiming));
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_tables_free, ctx->stream));
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_free, ctx->stream));
    for (int l = 0; l < ctx->levels; ++l)
    {
        kfb_intrinsics kl;
        kfb_level_intrinsics(intr, l, &kl);
        Level &L = ctx->L[l];
        L.k.w = kl.width; L.k.h = kl.height; L.k.fx = kl.fx; L.k.fy = kl.fy; L.k.cx = kl.cx; L.k.cy = kl.cy;
        const size_t n = (size_t)kl.width * kl.height;
        KFB_CUDA(ctx, cudaMalloc(&L.raw, n * sizeof(float)));
        KFB_CUDA(ctx, cudaMalloc(&L.depth, n * sizeof(float)));
        for (int f = 0; f < 2; ++f)
        {
            // vertex and normal map of a frame share one allocation (n = v + pixels): the cross-slab composite
            // reduces both with a single collective
            KFB_CUDA(ctx, cudaMalloc(&L.v[f], 2 * n * sizeof(float4)));
            L.n[f] = L.v[f] + n;
        }
    }
    // volume (or z-slab of it, with halo)
    const int Z = p->volu_dims[2];
    if (p->slab_z_end > p->slab_z_begin)
    {
        if (p->slab_z_begin < 0 || p->slab_z_end > Z) { ctx->err = "slab out of range"; return KFB_ERR_INVALID; }
        ctx->z0 = p->slab_z_begin - KFB_HALO < 0 ? 0 : p->slab_z_begin - KFB_HALO;
        ctx->z1 = p->slab_z_end + KFB_HALO > Z ? Z : p->slab_z_end + KFB_HALO;
    }
    else { ctx->z0 = 0; ctx->z1 = Z; }
    ctx->vol_logical = (size_t)p->volu_dims[0] * p->volu_dims[1] * (size_t)(ctx->z1 - ctx->z0);
    for (int i = 0; i < 3; ++i) ctx->voxel_size[i] = p->volu_range[i] / (float)p->volu_dims[i]; // tsdf_volume.cpp:16
    // brick grid of the stored planes: of the voxels (vol_index) and of the brick flags / distance map alike
    ctx->bdim[0] = (p->volu_dims[0] + 7) >> 3;
    ctx->bdim[1] = (p->volu_dims[1] + 7) >> 3;
    ctx->bz0 = ctx->z0 >> 3;
    ctx->bdim[2] = ((ctx->z1 - 1) >> 3) - ctx->bz0 + 1;
    ctx->vol_voxels = ((size_t)ctx->bdim[0] * ctx->bdim[1] * ctx->bdim[2]) << 9;
    KFB_CUDA(ctx, cudaMalloc(&ctx->vol, ctx->vol_voxels * sizeof(uint32_t)));
    const size_t n0 = (size_t)intr->width * intr->height;
    KFB_CUDA(ctx, cudaMalloc(&ctx->tab_thrz, n0 * sizeof(float2)));
    if (p->tsdf_max_weight < 1 || p->tsdf_max_weight > 32767) { ctx->err = "tsdf_max_weight must be in [1, 32767]"; return KFB_ERR_INVALID; }
    KFB_CUDA(ctx, cudaMalloc(&ctx->wtab, (size_t)(p->tsdf_max_weight + 1) * sizeof(float4)));
    KFB_CUDA(ctx, cudaMalloc(&ctx->zexit, sizeof(float)));
    {
        const size_t nb = (size_t)ctx->bdim[0] * ctx->bdim[1] * ctx->bdim[2];
        KFB_CUDA(ctx, cudaMalloc(&ctx->bricks, nb));
        KFB_CUDA(ctx, cudaMalloc(&ctx->bdist, nb));
        KFB_CUDA(ctx, cudaMalloc(&ctx->bdist_tmp, nb));
        KFB_CUDA(ctx, cudaMalloc(&ctx->bdist_tmp2, nb));
        KFB_CUDA(ctx, cudaMalloc(&ctx->bdirty, sizeof(int)));
    }
    KFB_CUDA(ctx, cudaMalloc(&ctx->tab4, n0 * sizeof(float4)));
    KFB_CUDA(ctx, cudaMalloc(&ctx->zsparse, 6 * n0 * sizeof(float2)));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->plan_hint_host, 64, cudaHostAllocDefault));
    memset(ctx->plan_hint_host, 0, 64);
    KFB_CUDA(ctx, cudaMalloc(&ctx->hit_t, n0 * sizeof(float)));
    KFB_CUDA(ctx, cudaMalloc(&ctx->shard_flag, 256));
    KFB_CUDA(ctx, cudaMemset(ctx->shard_flag, 0, 256));
    KFB_CUDA(ctx, cudaMalloc(&ctx->icp_partials, 1024 * 27 * sizeof(double)));
    KFB_CUDA(ctx, cudaMalloc(&ctx->icp_ticket, sizeof(unsigned int)));
    KFB_CUDA(ctx, cudaMemset(ctx->icp_ticket, 0, sizeof(unsigned int)));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->icp_host, sizeof(IcpHostResult), cudaHostAllocMapped));
    memset((void *)ctx->icp_host, 0, sizeof(IcpHostResult));
    KFB_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->icp_dev, (void *)ctx->icp_host, 0));
    memset(&ctx->icp_sched, 0, sizeof(ctx->icp_sched));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->dev_err_host, 64, cudaHostAllocMapped));
    memset((void *)ctx->dev_err_host, 0, 64);
    KFB_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->dev_err_dev, (void *)ctx->dev_err_host, 0));
    KFB_CUDA(ctx, cudaMalloc(&ctx->icp_devgate, sizeof(IcpDevGate)));
    KFB_CUDA(ctx, cudaMemset(ctx->icp_devgate, 0, sizeof(IcpDevGate)));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->icp_slots_host, sizeof(IcpHostSlot) * (KFB_ICP_MAX_ITERS + 1), cudaHostAllocMapped));
    memset((void *)ctx->icp_slots_host, 0, sizeof(IcpHostSlot) * (KFB_ICP_MAX_ITERS + 1));
    KFB_CUDA(ctx, cudaHostGetDevicePointer((void **)&ctx->icp_slots_dev, (void *)ctx->icp_slots_host, 0));
    ctx->icp_tagged_cap = ctx->sm_count;
    KFB_CUDA(ctx, cudaMalloc(&ctx->icp_tagged, sizeof(IcpTagged) * 2 * 27 * (size_t)ctx->icp_tagged_cap));
    KFB_CUDA(ctx, cudaMemset(ctx->icp_tagged, 0, sizeof(IcpTagged) * 2 * 27 * (size_t)ctx->icp_tagged_cap));
    KFB_CUDA(ctx, cudaMalloc(&ctx->counters, 8 * sizeof(unsigned long long)));
    KFB_CUDA(ctx, cudaMemset(ctx->counters, 0, 8 * sizeof(unsigned long long)));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->counters_host, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->pinned_depth, n0 * sizeof(float), cudaHostAllocDefault));
    KFB_CUDA(ctx, cudaMalloc(&ctx->depth_u16, n0 * sizeof(uint16_t)));
    KFB_CUDA(ctx, cudaMalloc(&ctx->render_dev, n0 * 3));
    KFB_CUDA(ctx, cudaHostAlloc((void **)&ctx->render_host, n0 * 3, cudaHostAllocDefault));
    for (int i = 0; i < 64; ++i) KFB_CUDA(ctx, cudaEventCreate(&ctx->events[i]));
    int rc = kfb_reset_frames(ctx);
    if (rc) return rc;
    rc = launch_build_wtab(ctx);
    if (rc) return rc;
    rc = kfb_reset_volume(ctx);
    if (rc) return rc;
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KFB_OK;
}

static void shard_close_peers(kfb_ctx *ctx);
void kfb_destroy(kfb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->fstream) cudaStreamSynchronize(ctx->fstream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->istream) cudaStreamSynchronize(ctx->istream);
    if (ctx->ostream) cudaStreamSynchronize(ctx->ostream);
    if (ctx->ev_ifork) cudaEventDestroy(ctx->ev_ifork);
    if (ctx->ev_ijoin) cudaEventDestroy(ctx->ev_ijoin);
    if (ctx->ev_plan_clean) cudaEventDestroy(ctx->ev_plan_clean);
    if (ctx->ev_sweep_main) cudaEventDestroy(ctx->ev_sweep_main);
    if (ctx->istream) cudaStreamDestroy(ctx->istream);
    if (ctx->ostream) cudaStreamDestroy(ctx->ostream);
    if (ctx->ev_upload) cudaEventDestroy(ctx->ev_upload);
    if (ctx->ev_front) cudaEventDestroy(ctx->ev_front);
    if (ctx->ev_free) cudaEventDestroy(ctx->ev_free);
    if (ctx->ev_tables_free) cudaEventDestroy(ctx->ev_tables_free);
    if (ctx->fstream) cudaStreamDestroy(ctx->fstream);
    for (int l = 0; l < KFB_MAX_LEVELS; ++l)
    {
        Level &L = ctx->L[l];
        if (L.raw) cudaFree(L.raw);
        if (L.depth) cudaFree(L.depth);
        for (int f = 0; f < 2; ++f) { if (L.v[f]) cudaFree(L.v[f]); }
    }
    if (ctx->vol) cudaFree(ctx->vol);
    if (ctx->tab_thrz) cudaFree(ctx->tab_thrz);
    if (ctx->wtab) cudaFree(ctx->wtab);
    if (ctx->zexit) cudaFree(ctx->zexit);
    if (ctx->zsparse) cudaFree(ctx->zsparse);
    if (ctx->bricks) cudaFree(ctx->bricks);
    if (ctx->bdist) cudaFree(ctx->bdist);
    if (ctx->bdist_tmp) cudaFree(ctx->bdist_tmp);
    if (ctx->bdist_tmp2) cudaFree(ctx->bdist_tmp2);
    if (ctx->bdirty) cudaFree(ctx->bdirty);
    if (ctx->tab4) cudaFree(ctx->tab4);
    if (ctx->plan_buf) cudaFree(ctx->plan_buf);
    if (ctx->plan_hint_host) cudaFreeHost(ctx->plan_hint_host);
    if (ctx->hit_t) cudaFree(ctx->hit_t);
    if (ctx->ray_cost) cudaFree(ctx->ray_cost);
    if (ctx->ev_ray_done) cudaEventDestroy(ctx->ev_ray_done);
    if (ctx->ev_ray_order) cudaEventDestroy(ctx->ev_ray_order);
    if (ctx->ev_order_gate) cudaEventDestroy(ctx->ev_order_gate);
    shard_close_peers(ctx);
    if (ctx->shard_flag) cudaFree(ctx->shard_flag);
    if (ctx->stage_keys) cudaFree(ctx->stage_keys);
    if (ctx->stage_maps) cudaFree(ctx->stage_maps);
    if (ctx->icp_partials) cudaFree(ctx->icp_partials);
    if (ctx->icp_ticket) cudaFree(ctx->icp_ticket);
    if (ctx->icp_host) cudaFreeHost((void *)ctx->icp_host);
    if (ctx->dev_err_host) cudaFreeHost((void *)ctx->dev_err_host);
    if (ctx->icp_devgate) cudaFree(ctx->icp_devgate);
    if (ctx->icp_slots_host) cudaFreeHost((void *)ctx->icp_slots_host);
    if (ctx->icp_tagged) cudaFree(ctx->icp_tagged);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->counters_host) cudaFreeHost(ctx->counters_host);
    if (ctx->pinned_depth) cudaFreeHost(ctx->pinned_depth);
    if (ctx->depth_u16) cudaFree(ctx->depth_u16);
    if (ctx->render_dev) cudaFree(ctx->render_dev);
    if (ctx->render_host) cudaFreeHost(ctx->render_host);
    if (ctx->cloud) cudaFree(ctx->cloud);
    for (int i = 0; i < 64; ++i) if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
    if (ctx->stream && ctx->own_stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
}

const char *kfb_last_error_string(const kfb_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

#define KFB_JOIN(ctx) do { const int _rc = join_front(ctx); if (_rc) return _rc; } while (0)

int kfb_synchronize(kfb_ctx *ctx)
{
    KFB_JOIN(ctx);
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->fstream));
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_device_error(ctx);
}

int kfb_set_stream(kfb_ctx *ctx, void *stream)
{
    if (ctx->icp_sched.active) { ctx->err = "kfb_set_stream inside an ICP schedule"; return KFB_ERR_INVALID; }
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->fstream));
    ctx->front_pending = 0;
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)stream;
    ctx->own_stream = 0;
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_tables_free, ctx->stream));
    return mark_free(ctx);
}

int kfb_reset_volume(kfb_ctx *ctx) { return launch_reset_volume(ctx); }

int kfb_reset_frames(kfb_ctx *ctx)
{
    KFB_JOIN(ctx);
    for (int l = 0; l < ctx->levels; ++l)
    {
        Level &L = ctx->L[l];
        const size_t n = (size_t)L.k.w * L.k.h;
        KFB_CUDA(ctx, cudaMemsetAsync(L.raw, 0, n * sizeof(float), ctx->stream));
        KFB_CUDA(ctx, cudaMemsetAsync(L.depth, 0, n * sizeof(float), ctx->stream));
        for (int f = 0; f < 2; ++f)
        {
            KFB_CUDA(ctx, cudaMemsetAsync(L.v[f], 0, n * sizeof(float4), ctx->stream));
            KFB_CUDA(ctx, cudaMemsetAsync(L.n[f], 0, n * sizeof(float4), ctx->stream));
        }
    }
    if (ctx->zsparse)
    {
        const int rc = launch_build_tables(ctx, ctx->stream); // tables of the (now empty) depth image
        if (rc) return rc;
    }
    return mark_free(ctx);
}

int kfb_upload_depth_mm(kfb_ctx *ctx, const float *host, int width, int height)
{
    if (!host || width != ctx->intr.width || height != ctx->intr.height) { ctx->err = "depth size mismatch"; return KFB_ERR_INVALID; }
    const size_t bytes = (size_t)width * height * sizeof(float);
    // pinned host memory goes straight over; device memory is copied D2D (frames already resident in
    // HBM); pageable memory is staged through the context's pinned buffer
    cudaPointerAttributes at;
    bool direct = false;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess)
        direct = (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    else cudaGetLastError();
    if (!direct)
    {
        KFB_CUDA(ctx, cudaStreamSynchronize(ctx->fstream)); // staging buffer may still be in flight
        memcpy(ctx->pinned_depth, host, bytes);
        host = ctx->pinned_depth;
    }
    const int rc = fork_front(ctx);
    if (rc) return rc;
    KFB_CUDA(ctx, cudaMemcpyAsync(ctx->L[0].raw, host, bytes, cudaMemcpyDefault, ctx->fstream));
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_upload, ctx->fstream));
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_front, ctx->fstream));
    ctx->front_pending = 1;
    return KFB_OK;
}

int kfb_upload_depth_mm_u16(kfb_ctx *ctx, const uint16_t *host, int width, int height)
{
    if (!host || width != ctx->intr.width || height != ctx->intr.height) { ctx->err = "depth size mismatch"; return KFB_ERR_INVALID; }
    const size_t n = (size_t)width * height, bytes = n * sizeof(uint16_t);
    cudaPointerAttributes at;
    bool direct = false;
    if (cudaPointerGetAttributes(&at, host) == cudaSuccess)
        direct = (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
    else cudaGetLastError();
    if (!direct)
    {
        KFB_CUDA(ctx, cudaStreamSynchronize(ctx->fstream)); // staging buffer may still be in flight
        memcpy(ctx->pinned_depth, host, bytes);             // the f32 staging buffer is large enough
        host = reinterpret_cast<const uint16_t *>(ctx->pinned_depth);
    }
    int rc = fork_front(ctx);
    if (rc) return rc;
    KFB_CUDA(ctx, cudaMemcpyAsync(ctx->depth_u16, host, bytes, cudaMemcpyDefault, ctx->fstream));
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_upload, ctx->fstream));
    rc = launch_u16_to_f32(ctx, ctx->depth_u16, ctx->L[0].raw, n, ctx->fstream);
    if (rc) return rc;
    KFB_CUDA(ctx, cudaEventRecord(ctx->ev_front, ctx->fstream));
    ctx->front_pending = 1;
    return KFB_OK;
}

int kfb_upload_wait(kfb_ctx *ctx)
{
    KFB_CUDA(ctx, cudaEventSynchronize(ctx->ev_upload));
    return KFB_OK;
}

int kfb_frontend(kfb_ctx *ctx) { return launch_frontend(ctx); }

int kfb_swap_frames(kfb_ctx *ctx)
{
    KFB_JOIN(ctx);
    ctx->pyramid_fresh = 0;
    const int t = ctx->cur; ctx->cur = ctx->prev; ctx->prev = t;
    return KFB_OK;
}

int kfb_icp_accumulate(kfb_ctx *ctx, int level, const float pose12[12], double out27[27])
{
    if (!pose12 || !out27) return KFB_ERR_INVALID;
    KFB_JOIN(ctx);
    return launch_icp(ctx, level, pose12, out27);
}

int kfb_icp_begin(kfb_ctx *ctx, const int iters_per_level[KFB_MAX_LEVELS])
{
    if (!iters_per_level) return KFB_ERR_INVALID;
    return icp_begin(ctx, iters_per_level);
}
int kfb_icp_step(kfb_ctx *ctx, const float pose12[12], double out27[27])
{
    if (!pose12 || !out27) return KFB_ERR_INVALID;
    KFB_JOIN(ctx);
    return icp_step(ctx, pose12, out27);
}
int kfb_icp_end(kfb_ctx *ctx) { return icp_end(ctx); }

int kfb_integrate(kfb_ctx *ctx, const float vol2cam12[12], uint64_t *n_updated)
{
    if (!vol2cam12) return KFB_ERR_INVALID;
    KFB_JOIN(ctx);
    return launch_integrate(ctx, vol2cam12, n_updated);
}

int kfb_integrate_plane_histogram(kfb_ctx *ctx, const float vol2cam12[12], uint32_t *host_hist)
{
    if (!vol2cam12 || !host_hist) return KFB_ERR_INVALID;
    KFB_JOIN(ctx);
    return launch_plane_histogram(ctx, vol2cam12, host_hist);
}

int kfb_raycast(kfb_ctx *ctx, const float cam2vol12[12], const float rinv9[9])
{
    if (!cam2vol12 || !rinv9) return KFB_ERR_INVALID;
    return launch_raycast(ctx, cam2vol12, rinv9);
}

int kfb_model_pyramid(kfb_ctx *ctx) { return launch_model_pyramid(ctx); }

// Buffers a slab context lets its peers write into (push composite): 0 = event keys [16][P] float,
// 1 = vertex + normal maps [16][2 P] float4 (slot r is written by rank r's raycast), 2 and 3 = "slab done" counters
// [16] (slot r written by rank r).  Only rank 0's copies are ever written, but every rank allocates and exports the
// same set so that the handle table is symmetric.  Allocated on first export.
#define KFB_SHARD_MAX 16
static int shard_alloc_stage(kfb_ctx *ctx)
{
    if (ctx->stage_keys) return KFB_OK;
    const size_t P = (size_t)ctx->intr.width * ctx->intr.height;
    KFB_CUDA(ctx, cudaMalloc(&ctx->stage_keys, KFB_SHARD_MAX * P * sizeof(float)));
    KFB_CUDA(ctx, cudaMalloc(&ctx->stage_maps, KFB_SHARD_MAX * 2 * P * sizeof(float4)));
    return KFB_OK;
}
static void *shard_export_ptr(kfb_ctx *ctx, int which)
{
    switch (which)
    {
    case 0: return ctx->stage_keys;
    case 1: return ctx->stage_maps;
    case 2: return ctx->shard_flag;
    case 3: return ctx->shard_flag;
    default: return nullptr;
    }
}
int kfb_ipc_export(kfb_ctx *ctx, int which, void *handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == KFB_IPC_HANDLE_BYTES, "IPC handle size");
    if (const int rcs = shard_alloc_stage(ctx)) return rcs;
    void *p = shard_export_ptr(ctx, which);
    if (!p || !handle64) return KFB_ERR_INVALID;
    cudaIpcMemHandle_t h;
    KFB_CUDA(ctx, cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, sizeof(h));
    return KFB_OK;
}
int kfb_shard_attach(kfb_ctx *ctx, int rank, int world, const void *handles)
{
    if (world < 2 || world > 16 || rank < 0 || rank >= world || !handles) return KFB_ERR_INVALID;
    if (ctx->shard_world) { ctx->err = "already attached"; return KFB_ERR_INVALID; }
    if (const int rcs = shard_alloc_stage(ctx)) return rcs;
    const unsigned char *hb = (const unsigned char *)handles;
    for (int r = 0; r < world; ++r)
    {
        void *ptr[4];
        for (int w = 0; w < 3; ++w) // (the fourth handle repeats the third)
        {
            if (r == rank) { ptr[w] = shard_export_ptr(ctx, w); continue; }
            cudaIpcMemHandle_t h;
            memcpy(&h, hb + ((size_t)r * 4 + w) * KFB_IPC_HANDLE_BYTES, sizeof(h));
            KFB_CUDA(ctx, cudaIpcOpenMemHandle(&ptr[w], h, cudaIpcMemLazyEnablePeerAccess));
        }
        ctx->peer_keys[r] = ptr[0]; ctx->peer_maps[0][r] = ptr[1]; ctx->peer_maps[1][r] = nullptr; ctx->peer_flag[r] = ptr[2];
    }
    ctx->shard_rank = rank; ctx->shard_world = world; ctx->shard_seq = 0;
    return KFB_OK;
}
int kfb_shard_attached(const kfb_ctx *ctx) { return ctx->shard_world > 0; }
static void shard_close_peers(kfb_ctx *ctx)
{
    for (int r = 0; r < ctx->shard_world && r < 16; ++r)
        if (r != ctx->shard_rank)
        {
            if (ctx->peer_keys[r]) cudaIpcCloseMemHandle(ctx->peer_keys[r]);
            if (ctx->peer_maps[0][r]) cudaIpcCloseMemHandle(ctx->peer_maps[0][r]);
            if (ctx->peer_maps[1][r]) cudaIpcCloseMemHandle(ctx->peer_maps[1][r]);
            if (ctx->peer_flag[r]) cudaIpcCloseMemHandle(ctx->peer_flag[r]);
        }
    memset(ctx->peer_keys, 0, sizeof(ctx->peer_keys)); memset(ctx->peer_maps, 0, sizeof(ctx->peer_maps)); memset(ctx->peer_flag, 0, sizeof(ctx->peer_flag));
    ctx->shard_world = 0;
}
int kfb_shard_detach(kfb_ctx *ctx)
{
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    shard_close_peers(ctx);
    cudaGetLastError();
    return KFB_OK;
}
int kfb_shard_composite(kfb_ctx *ctx)
{
    if (!ctx->shard_world) { ctx->err = "kfb_shard_composite without kfb_shard_attach"; return KFB_ERR_INVALID; }
    return launch_shard_composite(ctx);
}

int kfb_composite_mask(kfb_ctx *ctx, const float *min_key_dev)
{
    if (!min_key_dev) return KFB_ERR_INVALID;
    return launch_composite_mask(ctx, min_key_dev);
}

int kfb_extract_points(kfb_ctx *ctx, const float volpose12[12], float *host_points3, size_t cap, size_t *n_points)
{
    if (!volpose12 || !host_points3 || !n_points || cap == 0) return KFB_ERR_INVALID;
    return launch_extract(ctx, volpose12, host_points3, cap, n_points);
}

int kfb_render_phong(kfb_ctx *ctx, const float eye3[3], uint8_t *host_bgr)
{
    if (!eye3 || !host_bgr) return KFB_ERR_INVALID;
    return launch_render(ctx, 1, eye3, host_bgr);
}
int kfb_render_normals(kfb_ctx *ctx, uint8_t *host_bgr)
{
    if (!host_bgr) return KFB_ERR_INVALID;
    const float z[3] = {0, 0, 0};
    return launch_render(ctx, 0, z, host_bgr);
}

// ---- hooks ----------------------------------------------------------------------------------------
static int check_level(kfb_ctx *ctx, int level)
{
    if (level < 0 || level >= ctx->levels) { ctx->err = "level out of range"; return KFB_ERR_INVALID; }
    return KFB_OK;
}

int kfb_download_depth(kfb_ctx *ctx, int level, float *host)
{
    KFB_JOIN(ctx);
    if (check_level(ctx, level)) return KFB_ERR_INVALID;
    const Level &L = ctx->L[level];
    KFB_CUDA(ctx, cudaMemcpyAsync(host, L.depth, (size_t)L.k.w * L.k.h * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KFB_OK;
}
int kfb_download_raw_depth(kfb_ctx *ctx, int level, float *host)
{
    KFB_JOIN(ctx);
    if (check_level(ctx, level)) return KFB_ERR_INVALID;
    const Level &L = ctx->L[level];
    KFB_CUDA(ctx, cudaMemcpyAsync(host, L.raw, (size_t)L.k.w * L.k.h * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KFB_OK;
}
int kfb_upload_depth_m(kfb_ctx *ctx, int level, const float *host)
{
    KFB_JOIN(ctx);
    if (check_level(ctx, level)) return KFB_ERR_INVALID;
    const Level &L = ctx->L[level];
    KFB_CUDA(ctx, cudaMemcpyAsync(L.depth, host, (size_t)L.k.w * L.k.h * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (level == 0)
    {
        const int rc = launch_build_tables(ctx, ctx->stream);
        if (rc) return rc;
    }
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KFB_OK;
}
int kfb_download_maps(kfb_ctx *ctx, int frame, int level, float *host_v3, float *host_n3)
{
    KFB_JOIN(ctx);
    if (check_level(ctx, level)) return KFB_ERR_INVALID;
    const Level &L = ctx->L[level];
    const int f = frame == KFB_FRAME_CUR ? ctx->cur : ctx->prev;
    const size_t n = (size_t)L.k.w * L.k.h;
    float *tmp = nullptr;
    KFB_CUDA(ctx, cudaMalloc(&tmp, n * 3 * sizeof(float)));
    int rc = KFB_OK;
    for (int m = 0; m < 2 && rc == KFB_OK; ++m)
    {
        float *dst = m == 0 ? host_v3 : host_n3;
        if (!dst) continue;
        rc = launch_map_convert(ctx, m == 0 ? L.v[f] : L.n[f], tmp, n);
        if (rc == KFB_OK && cudaMemcpyAsync(dst, tmp, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = KFB_ERR_CUDA;
        if (rc == KFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = KFB_ERR_CUDA;
    }
    cudaFree(tmp);
    if (rc == KFB_ERR_CUDA && ctx->err.empty()) ctx->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}
int kfb_upload_maps(kfb_ctx *ctx, int frame, int level, const float *host_v3, const float *host_n3)
{
    KFB_JOIN(ctx);
    ctx->pyramid_fresh = 0;
    if (check_level(ctx, level)) return KFB_ERR_INVALID;
    const Level &L = ctx->L[level];
    const int f = frame == KFB_FRAME_CUR ? ctx->cur : ctx->prev;
    const size_t n = (size_t)L.k.w * L.k.h;
    float *tmp = nullptr;
    KFB_CUDA(ctx, cudaMalloc(&tmp, n * 3 * sizeof(float)));
    int rc = KFB_OK;
    for (int m = 0; m < 2 && rc == KFB_OK; ++m)
    {
        const float *src = m == 0 ? host_v3 : host_n3;
        if (!src) continue;
        if (cudaMemcpyAsync(tmp, src, n * 3 * sizeof(float), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) rc = KFB_ERR_CUDA;
        if (rc == KFB_OK) rc = launch_map_convert_in(ctx, tmp, m == 0 ? L.v[f] : L.n[f], n);
        if (rc == KFB_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = KFB_ERR_CUDA;
    }
    cudaFree(tmp);
    if (rc == KFB_ERR_CUDA && ctx->err.empty()) ctx->err = cudaGetErrorString(cudaGetLastError());
    return rc;
}
int kfb_download_volume(kfb_ctx *ctx, int16_t *host) { return launch_volume_copy(ctx, host, 0); }
int kfb_upload_volume(kfb_ctx *ctx, const int16_t *host)
{
    int rc = launch_volume_copy(ctx, const_cast<int16_t *>(host), 1);
    if (rc) return rc;
    rc = launch_rebuild_bricks(ctx);
    if (rc) return rc;
    KFB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KFB_OK;
}
size_t kfb_volume_voxels(const kfb_ctx *ctx) { return ctx->vol_logical; }

// ---- measurement ------------------------------------------------------------------------------------
int kfb_event_record(kfb_ctx *ctx, int slot)
{
    KFB_JOIN(ctx);
    if (slot < 0 || slot >= 64) return KFB_ERR_INVALID;
    KFB_CUDA(ctx, cudaEventRecord(ctx->events[slot], ctx->stream));
    return KFB_OK;
}
int kfb_event_elapsed_ms(kfb_ctx *ctx, int a, int b, float *ms)
{
    if (a < 0 || a >= 64 || b < 0 || b >= 64 || !ms) return KFB_ERR_INVALID;
    KFB_CUDA(ctx, cudaEventSynchronize(ctx->events[b]));
    KFB_CUDA(ctx, cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]));
    return KFB_OK;
}
int kfb_set_profiling(kfb_ctx *ctx, int on) { ctx->profiling = on; return KFB_OK; }
uint64_t kfb_launch_count(const kfb_ctx *ctx) { return ctx->launches; }
void *kfb_device_ptr(kfb_ctx *ctx, int which)
{
    switch (which)
    {
    case 0: return ctx->vol;
    case 1: return ctx->L[0].v[ctx->prev];
    case 2: return ctx->L[0].n[ctx->prev];
    case 3: return ctx->L[0].depth;
    case 4: return ctx->hit_t;
    default: return nullptr;
    }
}
void *kfb_stream(kfb_ctx *ctx) { return (void *)ctx->stream; }
void kfb_debug_icp_ring(kfb_ctx *ctx, uint64_t out128[128])
{
    for (int i = 0; i < 128; ++i) out128[i] = ctx->icp_host->post_ns[i / 4][i % 4];
}
uint64_t kfb_icp_fallback_count(const kfb_ctx *ctx) { return ctx ? ctx->icp_fallbacks : 0; }
uint64_t kfb_icp_mispredict_count(const kfb_ctx *ctx) { return ctx ? ctx->icp_mispredicts : 0; }
void kfb_debug_integrate_counts(kfb_ctx *ctx, uint64_t out6[6])
{
    // valid after a counting kfb_integrate call (n_updated != NULL), which synchronises
    out6[0] = ctx->counters_host[0]; out6[1] = ctx->counters_host[2]; out6[2] = ctx->counters_host[3];
    const unsigned int *pc = reinterpret_cast<const unsigned int *>(ctx->counters_host + 4);
    out6[3] = pc[0]; out6[4] = pc[1]; out6[5] = 0;
}

} // extern "C"
