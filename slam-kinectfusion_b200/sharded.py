"""z-slab sharded volumes over the GPUs of one NVLink box: one process per GPU,
`torch.distributed` for the plumbing.  Only the volume partitions: rank g stores and integrates planes
[g*Z/G, (g+1)*Z/G) (+ a 3-plane halo integrated redundantly, no exchange), marches the ray samples whose
voxel lies in its slab, and the first terminal ray event over all slabs is selected with two collectives:

  (default) ONE kernel on rank 0 over NVLink peer memory (CUDA IPC, kfb_shard_attach / kfb_shard_composite): it
    waits for every slab's "done" flag, reads the peers' event keys, and pulls the winner's vertex and normal;
  (KFB_COMPOSITE_NCCL=1) the same selection with two collectives:
    all_reduce(MIN) over the per-pixel event keys (ray length of the slab's first hit / back-face stop)
    kfb_composite_mask: every rank zeroes its vertex / normal maps where it does not hold the winning key
    reduce(SUM, int32 view) of the maps to rank 0 (x + 0 == x exactly in integers, so the composite is
    bit-identical to the single-GPU raycast)

ICP stays on rank 0 (it owns the composited model maps); only {tracking_ok, 4x3 pose} is handed to the other
ranks (shared-memory mailbox on the node, or torch.distributed broadcast with KFB_POSE_NCCL=1).
The frame logic itself is the C++ facade's (kf::kinectfusion with kf::ShardComm callbacks); this module
supplies the two collectives and the launch/bench glue.  `composite` and `broadcast_pose` are written
against plain tensors so the same code runs over gloo on CPU tensors in tests/test_sharded.py.
"""
import ctypes as C
import json
import os
import time

import numpy as np

HALO = 3  # planes, must match KFB_HALO in csrc/kfb_api.cu


def slab_range(Z, world, rank):
    """Owned planes [zb, ze) of rank `rank`: contiguous, covering [0, Z) exactly once."""
    return (rank * Z) // world, ((rank + 1) * Z) // world


def stored_range(Z, world, rank, bounds=None):
    zb, ze = (bounds[rank], bounds[rank + 1]) if bounds is not None else slab_range(Z, world, rank)
    return max(zb - HALO, 0), min(ze + HALO, Z)


def balanced_bounds(hist, world, min_planes=8):
    """Slab boundaries b[0] = 0 < b[1] < ... < b[world] = Z such that every slab carries about the same
    integration work.  hist[z] = visited 4-voxel groups on plane z (kfb_integrate_plane_histogram); a small
    constant per plane stands for the per-plane overheads.  Deterministic: every rank computes the same."""
    h = np.asarray(hist, np.float64)
    Z = len(h)
    work = h + 0.02 * max(float(h.max()), 1.0)
    cum = np.concatenate([[0.0], np.cumsum(work)])
    b = [0]
    for r in range(1, world):
        z = int(np.searchsorted(cum, cum[-1] * r / world))
        z = max(z, b[-1] + min_planes)
        z = min(z, Z - (world - r) * min_planes)
        b.append(z)
    b.append(Z)
    return b


def measure_plane_histogram(K, hp, depth_mm, device):
    """Work histogram of the first frame at the bootstrap pose (identity), from a throw-away context that stores
    only a handful of planes."""
    from . import binding
    p = binding.default_params(int(hp.volu_dims[0]))
    for i in range(3):
        p.volu_dims[i] = hp.volu_dims[i]
        p.volu_range[i] = hp.volu_range[i]
    p.volu_trun_dist = hp.volu_trun_dist
    p.slab_z_begin, p.slab_z_end = 1, 2
    ctx = binding.Context(K, p, device=device)
    ctx.upload_depth_mm(depth_mm)
    ctx.frontend()
    hist = ctx.plane_histogram(np.array(hp.volu_pose, np.float32))   # vol2cam at the identity camera pose = volume pose
    ctx.close()
    return hist


def broadcast_pose(dist, msg13, device):
    """msg13: numpy float32[13] (valid on rank 0).  Returns the broadcast copy (numpy)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(msg13, np.float32)).to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def composite(dist, keys, maps_i32, mask_fn, dst=0, scratch=None):
    """keys: float32 tensor [P]; maps_i32: int32 view of the slab's model vertex+normal maps (one contiguous
    buffer); mask_fn(min_keys) zeroes the maps where this rank does not hold the winning key.  After the call
    rank `dst` holds the composite in its maps (in place)."""
    min_keys = scratch if scratch is not None else keys.clone()
    if scratch is not None:
        min_keys.copy_(keys)
    dist.all_reduce(min_keys, op=dist.ReduceOp.MIN)
    mask_fn(min_keys)
    dist.reduce(maps_i32, dst=dst, op=dist.ReduceOp.SUM)
    return min_keys


class PoseMailbox:
    """{tracking_ok, pose12} from rank 0 to the other ranks of the same node through POSIX shared memory (52
    bytes + a sequence number; all ranks of a sharded volume sit on one NVLink box).  This is launcher
    plumbing, not a data-path collective; `broadcast_pose` over torch.distributed is the portable equivalent."""

    def __init__(self, dist, rank, tag, world=None):
        from multiprocessing import shared_memory
        self.rank = rank
        self.world = world if world is not None else dist.get_world_size()
        name = f"kfb_pose_{tag}"
        size = 1024
        if rank == 0:
            try:
                old = shared_memory.SharedMemory(name=name)
                old.close()
                old.unlink()
            except FileNotFoundError:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=size)
            self.shm.buf[:size] = bytes(size)
        dist.barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
        self.seq_view = np.ndarray((1,), np.int64, self.shm.buf, 0)
        self.ack_view = np.ndarray((64,), np.int64, self.shm.buf, 64)     # one acknowledgement counter per rank
        self.msg_view = np.ndarray((13,), np.float32, self.shm.buf, 640)
        self.seq = 0
        dist.barrier()

    def _spin(self, cond):
        spins = 0
        while not cond():
            spins += 1
            if spins > 100_000_000:
                raise TimeoutError("pose mailbox: peer never arrived")

    def exchange(self, msg13):
        """msg13: numpy float32[13], valid on rank 0 on entry, on every rank on return."""
        self.seq += 1
        if self.rank == 0:
            # every reader has taken the previous message before it is overwritten
            self._spin(lambda: all(self.ack_view[r] >= self.seq - 1 for r in range(1, self.world)))
            self.msg_view[:] = msg13
            self.seq_view[0] = self.seq          # x86 TSO: payload before flag
        else:
            self._spin(lambda: self.seq_view[0] >= self.seq)
            msg13[:] = self.msg_view
            self.ack_view[self.rank] = self.seq

    def close(self):
        try:
            self.seq_view = self.msg_view = self.ack_view = None
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


class DevView:
    """Zero-copy torch view of a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def dev_tensor(ptr, shape, typestr, device):
    import torch
    return torch.as_tensor(DevView(ptr, shape, typestr), device=device)


class ShardedKinectFusion:
    """kf::kinectfusion on one z-slab of the volume; `dist` is an initialised torch.distributed (NCCL)."""

    def __init__(self, K, hp, dist, rank, world, local, first_depth=None):
        """first_depth: the first frame (mm).  When given, slab heights are balanced by the integration work it
        implies (balanced_bounds); otherwise the volume is cut into equal slabs."""
        import torch
        from . import host
        self.dist, self.rank, self.world = dist, rank, world
        self.device = torch.device("cuda", local)
        Z = hp.volu_dims[2]
        self.bounds = None
        if first_depth is not None and os.environ.get("KFB_SLABS_EQUAL") is None:
            self.bounds = balanced_bounds(measure_plane_histogram(K, hp, first_depth, local), world)
            hp.slab_z_begin, hp.slab_z_end = self.bounds[rank], self.bounds[rank + 1]
        else:
            hp.slab_z_begin, hp.slab_z_end = slab_range(Z, world, rank)
        hp.shard_rank, hp.shard_world = rank, world
        hp.device = local
        self.kf = host.KinectFusion(K, hp)
        self.ctx = self.kf.context()
        # all device work of the context and the collectives are ordered on torch's current stream
        self.ctx.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        self.P = K.width * K.height
        self.mailbox = PoseMailbox(dist, rank, os.environ.get("MASTER_PORT", "0")) if os.environ.get("KFB_POSE_NCCL") is None else None
        self.min_keys = torch.empty(self.P, dtype=torch.float32, device=self.device)
        self._view_cache = {}
        self.kf.set_shard_comm(self._bcast, self._composite)
        self.p2p = os.environ.get("KFB_COMPOSITE_NCCL") is None
        if self.p2p:
            # exchange CUDA IPC handles once; from then on the composite is one kernel over NVLink peer memory
            mine = torch.from_numpy(self.ctx.ipc_export()).to(self.device)
            allh = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allh, mine)
            self.ctx.shard_attach(rank, world, torch.stack(allh).cpu().numpy().reshape(-1))

    def _views(self):
        # the model maps swap buffers on the bootstrap frame: take the pointers per call (cheap)
        d = self.device
        pk, pm = self.ctx.device_ptr(4), self.ctx.device_ptr(1)
        if (pk, pm) not in self._view_cache:
            self._view_cache[(pk, pm)] = (dev_tensor(pk, (self.P,), "<f4", d),
                                          dev_tensor(pm, (self.P * 8,), "<i4", d))   # vertex map + normal map, contiguous
        return self._view_cache[(pk, pm)]

    def _bcast(self, p):
        try:
            msg = np.ctypeslib.as_array(p, shape=(13,))
            if self.mailbox is not None:
                self.mailbox.exchange(msg)
            else:
                msg[:] = broadcast_pose(self.dist, msg, self.device)
            return 0
        except Exception as e:  # noqa: BLE001 - reported through the C return code
            print("broadcast_pose failed:", e, flush=True)
            return 1

    def _composite(self):
        try:
            keys, maps = self._views()
            composite(self.dist, keys, maps, lambda mk: self.ctx.composite_mask(mk.data_ptr()), scratch=self.min_keys)
            return 0
        except Exception as e:  # noqa: BLE001
            print("composite failed:", e, flush=True)
            return 1

    def pipeline_ptr(self, ptr, w, h):
        return self.kf.pipeline_ptr(ptr, w, h)

    def close(self):
        if self.mailbox is not None:
            self.dist.barrier()
            self.mailbox.close()
            self.mailbox = None


def run_bench(args, dist, rank, world, local, dims, K, frames, host_pin, dev_frames, METRIC, UNIT, measured_peak_hbm,
              ClockSampler):
    """bench.py's N > 1 leg: every rank runs the sharded pipeline on the same frames; value = updated voxels
    per frame summed over the ranks' OWNED planes / device time per frame (max over ranks)."""
    import torch
    from . import host
    W, S = args.warmup, args.steps
    n_frames = len(frames)
    w, h = K.width, K.height
    hp = host.default_host_params(dims)
    skf = ShardedKinectFusion(K, hp, dist, rank, world, local, first_depth=frames[0][1])
    ctx = skf.ctx
    dev = torch.device("cuda", local)

    def run(ptrs):
        skf.kf.reset()
        t0 = None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        l0 = 0
        for i, p in enumerate(ptrs):
            if i == 1 + W:
                dist.barrier()
                torch.cuda.synchronize(dev)
                l0 = ctx.launch_count()
                e0.record()
                t0 = time.perf_counter()
            rc = skf.pipeline_ptr(p, w, h)
            if rc != 0:
                raise SystemExit(f"rank {rank}: pipeline rc={rc} at frame {i}: {skf.kf.last_error()}")
        e1.record()
        torch.cuda.synchronize(dev)
        wall = time.perf_counter() - t0
        dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1]), ctx.launch_count() - l0

    dptr = [dev_frames[i].data_ptr() for i in range(n_frames)]
    hptr = [host_pin[i].data_ptr() for i in range(n_frames)]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms, _, launches = run(dptr)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms, e2e_wall, _ = run(hptr)
    e2e_ms = max(e2e_ms, e2e_wall)
    poses = skf.kf.poses()

    # outside the timed region: updated voxels of this rank's stored planes, and the integrate kernel's duration
    ctx.set_profiling(True)
    volpose = np.array(hp.volu_pose, np.float32).reshape(3, 4)

    def vol2cam(p12):
        Pm = np.vstack([np.asarray(p12, np.float64).reshape(3, 4), [0, 0, 0, 1]])
        V = np.vstack([volpose.astype(np.float64), [0, 0, 0, 1]])
        return (np.linalg.inv(Pm) @ V)[:3].astype(np.float32).reshape(12)

    U, k_ms = [], []
    for i in range(1 + W, n_frames, max(1, S // 8)):
        ctx.upload_depth_mm_ptr(dptr[i], w, h)
        ctx.frontend()
        v2c = vol2cam(poses[i])
        U.append(ctx.integrate(v2c, count=True))
        for _ in range(2):
            ctx.integrate(v2c)
            k_ms.append(ctx.event_elapsed_ms(60, 61))
    ctx.set_profiling(False)
    zs0, zs1 = stored_range(dims, world, rank, skf.bounds)
    zb, ze = (skf.bounds[rank], skf.bounds[rank + 1]) if skf.bounds is not None else slab_range(dims, world, rank)
    own_frac = (ze - max(zb, 1)) / max(zs1 - max(zs0, 1), 1)   # halo planes are integrated redundantly: not counted
    stat = torch.tensor([float(np.mean(U)) * own_frac, float(np.mean(U)), float(np.mean(k_ms))], device=dev, dtype=torch.float64)
    gathered = [torch.zeros_like(stat) for _ in range(world)]
    dist.all_gather(gathered, stat)
    skf.close()
    if rank != 0:
        return
    U_owned = sum(float(g[0]) for g in gathered)
    k_max = max(float(g[2]) for g in gathered)
    U_stored = sum(float(g[1]) for g in gathered)
    peak, peak_src = measured_peak_hbm()
    achieved = 8.0 * U_stored / (k_max * 1e-3) / 1e9       # aggregate over the ranks, slowest rank's kernel time
    ms_per_frame = dev_ms / S
    line = {
        "metric": METRIC, "value": U_owned / (ms_per_frame * 1e-3), "unit": UNIT, "n_gpus": world, "steps": S, "warmup": W,
        "ms_per_step": ms_per_frame, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 -> int16 tsdf", "data": "synthetic",
        "config": {"workload": f"640x480 depth, {dims}^3 TSDF over 3 m sharded into {world} z-slabs (~512^3 voxels per GPU), "
                               "ICP 10/5/4 on rank 0, per-slab raycast + first-hit composite, 300-frame looped synthetic trajectory",
                   "l2": "inputs larger than L2: every rank sweeps its >= 512 MiB slab each frame",
                   "frames_timed": S, "updated_voxels_per_frame": U_owned, "swept_voxels_per_frame": dims * dims * (dims - 1),
                   "collectives_per_frame": ("pose mailbox 52 B (shared memory); composite = one kernel over NVLink peer memory" if skf.p2p else "pose mailbox 52 B (shared memory), all_reduce(min) 1.2 MB, reduce(sum) 9.8 MB")},
        "frame_device_ms": ms_per_frame,
        "slab_bounds": skf.bounds if skf.bounds is not None else [slab_range(dims, world, r)[0] for r in range(world)] + [dims],
        "final_pose": [float(x) for x in poses[-1]],
        "e2e": {"value": U_owned / (e2e_ms / S * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms / S,
                "h2d_bytes_per_step": w * h * 4 * world, "d2h_bytes_per_step": 19 * 27 * 16 + 52 * world,
                "api": "kf::kinectfusion::pipeline(depth_mm) on every rank (z-slab sharded, kf::ShardComm over NCCL), pinned host frames"},
        "gpu_launches": int(launches) * world,
        "roofline": {"bound": "hbm", "kernel": "integrate_kernel", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                     "frac": achieved / (peak * world), "traffic": None, "peak_source": peak_src + f" x {world} GPUs",
                     "kernel_ms": k_max, "algorithmic_bytes": 8.0 * U_stored},
        "clocks": clocks,
    }
    print(json.dumps(line))
