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

    def __init__(self, K, hp, dist, rank, world, local, first_depth=None, bounds=None):
        """first_depth: the first frame (mm).  When given, slab heights are balanced by the work it implies
        (balanced_bounds over kfb_integrate_plane_histogram); `bounds` (world + 1 plane numbers) overrides that
        (calibrated(): bounds corrected by measured per-rank times); otherwise the volume is cut into equal slabs."""
        import torch
        from . import host
        self.dist, self.rank, self.world = dist, rank, world
        self.device = torch.device("cuda", local)
        Z = hp.volu_dims[2]
        self.bounds = None
        if bounds is not None:
            self.bounds = [int(b) for b in bounds]
            hp.slab_z_begin, hp.slab_z_end = self.bounds[rank], self.bounds[rank + 1]
        elif first_depth is not None and os.environ.get("KFB_SLABS_EQUAL") is None:
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
        tag = os.environ.get("KFB_MAILBOX_TAG", os.environ.get("MASTER_PORT", "0"))
        # pose hand-off: the facade's native shared-memory mailbox (default; no Python on the per-frame path), the
        # Python one (KFB_POSE_PYMAILBOX=1) or a torch.distributed broadcast (KFB_POSE_NCCL=1)
        self.mailbox = None
        self.native_mailbox = None
        self.mailbox_us = []
        if os.environ.get("KFB_POSE_NCCL") is None:
            if os.environ.get("KFB_POSE_PYMAILBOX") is not None:
                self.mailbox = PoseMailbox(dist, rank, tag)
            else:
                name = f"/kfb_pose_{os.getuid()}_{tag}"
                if rank == 0:
                    self.native_mailbox = host.NativePoseMailbox(name, rank, world)
                dist.barrier()
                if rank != 0:
                    self.native_mailbox = host.NativePoseMailbox(name, rank, world)
                dist.barrier()
        self.min_keys = torch.empty(self.P, dtype=torch.float32, device=self.device)
        self._view_cache = {}
        if self.native_mailbox is not None:
            self.kf.set_pose_mailbox(self.native_mailbox, self._composite)
        else:
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
            t0 = time.perf_counter()
            msg = np.ctypeslib.as_array(p, shape=(13,))
            if self.mailbox is not None:
                self.mailbox.exchange(msg)
            else:
                msg[:] = broadcast_pose(self.dist, msg, self.device)
            self.mailbox_us.append((time.perf_counter() - t0) * 1e6)
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
        if self.p2p:
            # every rank unmaps its peers' buffers before anybody frees them
            self.ctx.shard_detach()
            self.p2p = False
        self.dist.barrier()
        if self.mailbox is not None:
            self.mailbox.close()
            self.mailbox = None
        if self.native_mailbox is not None:
            self.native_mailbox.close()
            self.native_mailbox = None


def _measure_config(args, dist, rank, world, local, dims, K, frames, host_pin, dev_frames, ClockSampler, with_e2e, tag):
    """One sharded configuration (a `dims`^3 volume cut into `world` z-slabs) on the given frames: device-timed
    frame loop (max over ranks), optionally the same from pinned host frames, the per-stage kernel durations in
    situ (slowest rank), updated-voxel counts, and -- on rank 0, after the slab contexts are gone -- the SAME
    frames through the single-GPU pipeline at the same dims: its poses must equal the sharded run's bit for bit
    (N-GPU == 1-GPU), and its frame time is the N = 1 point of the strong-scaling curve."""
    import torch
    from . import host
    W, S = args.warmup, args.steps
    n_frames = len(frames)
    w, h = K.width, K.height
    hp = host.default_host_params(dims)
    os.environ["KFB_MAILBOX_TAG"] = tag
    # slabs cut by the sweep's plan, then corrected by two measured rounds over the first frames (outside any timed region)
    skf = calibrated(K, lambda: host.default_host_params(dims), dist, rank, world, local, [f[1] for f in frames[:min(6, n_frames)]])
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
    e2e_ms = None
    if with_e2e:
        e2e_ms, e2e_wall, _ = run(hptr)
        e2e_ms = max(e2e_ms, e2e_wall)
    poses = skf.kf.poses()

    # ---- per-stage durations in situ: the same pipelined sequence with the library's profiling events on; every
    # 4th frame all ranks synchronise and read that frame's events (integrate kernel 60/61, whole integrate call
    # 56/57, slab raycast 58/59; rank 0: whole-schedule ICP kernel 54/55, composite kernel 52/53)
    ctx.set_profiling(True)
    skf.kf.reset()
    skf.mailbox_us = []
    st = {k: [] for k in ("integrate_kernel", "integrate_call", "raycast", "icp", "composite")}
    for i, p in enumerate(dptr):
        if skf.pipeline_ptr(p, w, h) != 0:
            raise SystemExit(f"rank {rank}: tracking failure at frame {i} (profiled pass)")
        if i > W and i % 4 == 0:
            ctx.synchronize()
            st["integrate_kernel"].append(ctx.event_elapsed_ms(60, 61))
            st["integrate_call"].append(ctx.event_elapsed_ms(56, 57))
            st["raycast"].append(ctx.event_elapsed_ms(58, 59))
            if rank == 0:
                for name, (a, b) in (("icp", (54, 55)), ("composite", (52, 53))):
                    try:
                        st[name].append(ctx.event_elapsed_ms(a, b))
                    except Exception:  # noqa: BLE001 - events never recorded (KFB_ICP_DIRECT, NCCL composite)
                        pass
    mailbox_us = float(np.mean(skf.mailbox_us[1 + W:])) if len(skf.mailbox_us) > 1 + W else 0.0
    # ---- updated voxels of this rank's stored planes (counting variant) at the tracked poses
    volpose = np.array(hp.volu_pose, np.float32).reshape(3, 4)

    def vol2cam(p12):
        Pm = np.vstack([np.asarray(p12, np.float64).reshape(3, 4), [0, 0, 0, 1]])
        V = np.vstack([volpose.astype(np.float64), [0, 0, 0, 1]])
        return (np.linalg.inv(Pm) @ V)[:3].astype(np.float32).reshape(12)

    U = []
    for i in range(1 + W, n_frames, max(1, S // 8)):
        ctx.upload_depth_mm_ptr(dptr[i], w, h)
        ctx.frontend()
        U.append(ctx.integrate(vol2cam(poses[i]), count=True))
    ctx.set_profiling(False)
    zs0, zs1 = stored_range(dims, world, rank, skf.bounds)
    zb, ze = (skf.bounds[rank], skf.bounds[rank + 1]) if skf.bounds is not None else slab_range(dims, world, rank)
    own_frac = (ze - max(zb, 1)) / max(zs1 - max(zs0, 1), 1)   # halo planes are integrated redundantly: not counted
    mean = lambda v: float(np.mean(v)) if len(v) else 0.0  # noqa: E731
    stat = torch.tensor([mean(U) * own_frac, mean(U), mean(st["integrate_kernel"]), mean(st["integrate_call"]), mean(st["raycast"])],
                        device=dev, dtype=torch.float64)
    gathered = [torch.zeros_like(stat) for _ in range(world)]
    dist.all_gather(gathered, stat)
    bounds = skf.bounds if skf.bounds is not None else [slab_range(dims, world, r)[0] for r in range(world)] + [dims]
    p2p = skf.p2p
    calibration = getattr(skf, "calibration", None)
    skf.close()
    skf.kf.close()
    del skf
    torch.cuda.synchronize(dev)
    # ---- rank 0: the same frames on ONE GPU at the same dims (parity + the N = 1 point of the strong curve)
    single = None
    if rank == 0:
        hp1 = host.default_host_params(dims)
        hp1.device = local
        kf1 = host.KinectFusion(K, hp1)
        c1 = kf1.context()
        c1.set_profiling(True)
        k1 = []
        for i, p in enumerate(dptr):
            if i == 1 + W:
                c1.synchronize()
                c1.event_record(0)
            if kf1.pipeline_ptr(p, w, h) != 0:
                raise SystemExit(f"single-GPU check run lost tracking at frame {i}")
        c1.event_record(1)
        c1.synchronize()
        k1.append(c1.event_elapsed_ms(60, 61))
        poses1 = kf1.poses()
        single = {"ms_per_step": c1.event_elapsed_ms(0, 1) / S, "integrate_kernel_ms_last_frame": k1[-1],
                  "poses_equal": bool(len(poses1) == len(poses) and np.array_equal(np.asarray(poses1), np.asarray(poses)))}
        kf1.close()
    dist.barrier()
    return {
        "dims": dims, "ms_per_step": dev_ms / S, "e2e_ms_per_step": (e2e_ms / S) if e2e_ms is not None else None,
        "launches": launches, "clocks": clocks, "poses": poses, "bounds": bounds, "p2p": p2p, "calibration": calibration,
        "U_owned": sum(float(g[0]) for g in gathered), "U_stored": sum(float(g[1]) for g in gathered),
        "stage_ms": {"integrate_kernel_slowest_rank": max(float(g[2]) for g in gathered),
                     "integrate_call_slowest_rank": max(float(g[3]) for g in gathered),
                     "raycast_slowest_rank": max(float(g[4]) for g in gathered),
                     "integrate_kernel_per_rank": [float(g[2]) for g in gathered],
                     "raycast_per_rank": [float(g[4]) for g in gathered],
                     "icp_kernel_rank0": mean(st["icp"]), "composite_kernel_rank0": mean(st["composite"]),
                     "pose_mailbox_host_us_rank0": mailbox_us},
        "single": single,
    }


def rebalance(bounds, times, min_planes=8):
    """New slab bounds from measured per-rank times: the cost density is taken as constant inside each old slab and
    the total is cut into equal parts.  Deterministic (every rank computes the same from the gathered times)."""
    b = [int(x) for x in bounds]
    world = len(b) - 1
    Z = b[-1]
    dens = np.zeros(Z)
    for r in range(world):
        dens[b[r]:b[r + 1]] = max(float(times[r]), 1e-9) / max(b[r + 1] - b[r], 1)
    cum = np.concatenate([[0.0], np.cumsum(dens)])
    nb = [0]
    for r in range(1, world):
        z = int(np.searchsorted(cum, cum[-1] * r / world))
        z = max(z, nb[-1] + min_planes)
        z = min(z, Z - (world - r) * min_planes)
        nb.append(z)
    nb.append(Z)
    return nb


def calibrated(K, make_hp, dist, rank, world, local, frames_mm, rounds=3):
    """A ShardedKinectFusion whose slab bounds were corrected by measurement: the plan-based cut (cheap, no
    communication) balances the sweep but knows the raycast only by a model; here the first frames are run, every rank
    reports integrate call + raycast of the last one (the library's profiling events), and the slabs are re-cut half
    way towards the cut the measurement suggests (rebalance assumes a constant cost density inside a slab, which
    overshoots where the cost sits at one end of it).  After `rounds` corrections the best cut seen (smallest time of
    the slowest rank) is kept.  Volumes are rebuilt in between, so this belongs before the sequence starts; a running
    system would have to move planes between ranks instead (not built).  make_hp() -> fresh HostParams."""
    import torch
    skf = ShardedKinectFusion(K, make_hp(), dist, rank, world, local, first_depth=frames_mm[0])
    if os.environ.get("KFB_SLABS_NOCALIB") is not None or skf.bounds is None:
        return skf
    w, h = K.width, K.height
    dev = torch.device("cuda", local)
    gpu_frames = [torch.from_numpy(np.ascontiguousarray(f, np.float32)).to(dev) for f in frames_mm]

    def measure(s):
        s.ctx.set_profiling(True)
        for f in gpu_frames:
            if s.pipeline_ptr(f.data_ptr(), w, h) != 0:
                raise RuntimeError("calibration run lost tracking")
        s.ctx.synchronize()
        t = s.ctx.event_elapsed_ms(56, 57) + s.ctx.event_elapsed_ms(58, 59)
        s.ctx.set_profiling(False)
        tt = torch.tensor([t], device=dev, dtype=torch.float64)
        allt = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(allt, tt)
        return [float(x[0]) for x in allt]

    def rebuild(s, nb):
        s.close()
        s.kf.close()
        del s
        torch.cuda.synchronize(dev)
        return ShardedKinectFusion(K, make_hp(), dist, rank, world, local, bounds=nb)

    seen = []
    for r in range(rounds + 1):
        times = measure(skf)
        seen.append({"bounds": list(skf.bounds), "times_ms": times})
        if r == rounds:
            break
        target = rebalance(skf.bounds, times)
        nb = [int(round(0.5 * (a + b))) for a, b in zip(skf.bounds, target)]
        nb[0], nb[-1] = 0, skf.bounds[-1]
        if nb == list(skf.bounds):
            break
        skf = rebuild(skf, nb)
    best = min(seen, key=lambda c: max(c["times_ms"]))
    skf = rebuild(skf, best["bounds"])       # fresh volume for the sequence, at the best cut seen
    skf.calibration = {"rounds": seen, "chosen": best["bounds"]}
    return skf


def run_bench(args, dist, rank, world, local, dims, K, frames, host_pin, dev_frames, METRIC, UNIT, measured_peak_hbm,
              ClockSampler, workload_label=None):
    """bench.py's N > 1 leg.  The driver's line is the WEAK-scaled job (~512^3 voxels per GPU: every rank runs
    the sharded pipeline on the same frames; value = updated voxels per frame summed over the ranks' OWNED
    planes / device time per frame, max over ranks).  The same line carries, under "strong", BASELINE's large
    volume (--strong-dims, default 2048^3) on the same N GPUs next to the same volume on ONE GPU, and under
    "parity" the bit-for-bit comparison of the sharded poses with the single-GPU pipeline's."""
    w, h = K.width, K.height
    S, W = args.steps, args.warmup
    port = os.environ.get("MASTER_PORT", "0")
    main = _measure_config(args, dist, rank, world, local, dims, K, frames, host_pin, dev_frames, ClockSampler, True, f"{port}_w")
    strong = None
    sd = getattr(args, "strong_dims", 0) or 0
    if sd and sd != dims:
        strong = _measure_config(args, dist, rank, world, local, sd, K, frames, host_pin, dev_frames, ClockSampler, False, f"{port}_s")
    if rank != 0:
        return
    peak, peak_src = measured_peak_hbm()
    k_max = main["stage_ms"]["integrate_kernel_slowest_rank"]
    achieved = 8.0 * main["U_stored"] / (k_max * 1e-3) / 1e9       # aggregate over the ranks, slowest rank's kernel time
    ms_per_frame = main["ms_per_step"]
    label = workload_label(dims) if workload_label else f"640x480 depth, {dims}^3 TSDF over 3 m"
    parity = {"final_pose_equal": bool(main["single"]["poses_equal"]),
              "what": f"all {1 + W + S} poses of the {world}-GPU sharded run == the single-GPU pipeline's at {dims}^3, bit for bit"}
    line = {
        "metric": METRIC, "value": main["U_owned"] / (ms_per_frame * 1e-3), "unit": UNIT, "n_gpus": world, "steps": S, "warmup": W,
        "ms_per_step": ms_per_frame, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 -> int16 tsdf", "data": "synthetic",
        "config": {"workload": label,
                   "sharding": f"{world} z-slabs (~512^3 voxels per GPU), ICP on rank 0, per-slab raycast + first-hit composite",
                   "l2": "inputs larger than L2: every rank sweeps its >= 512 MiB slab each frame",
                   "frames_timed": S, "updated_voxels_per_frame": main["U_owned"], "swept_voxels_per_frame": dims * dims * (dims - 1),
                   "collectives_per_frame": ("pose mailbox 52 B (shared memory); composite = one kernel over NVLink peer memory" if main["p2p"] else "pose mailbox 52 B (shared memory), all_reduce(min) 1.2 MB, reduce(sum) 9.8 MB")},
        "frame_device_ms": ms_per_frame,
        "slab_bounds": main["bounds"], "slab_calibration": main["calibration"],
        "final_pose": [float(x) for x in main["poses"][-1]],
        "parity": parity,
        "stage_ms": main["stage_ms"],
        "single_gpu_same_dims": {k: v for k, v in main["single"].items()},
        "e2e": {"value": main["U_owned"] / (main["e2e_ms_per_step"] * 1e-3), "unit": UNIT, "ms_per_step": main["e2e_ms_per_step"],
                "h2d_bytes_per_step": w * h * 4 * world, "d2h_bytes_per_step": 19 * 27 * 16 + 52 * world,
                "api": "kf::kinectfusion::pipeline(depth_mm) on every rank (z-slab sharded, kf::ShardComm), pinned host frames"},
        "gpu_launches": int(main["launches"]) * world,
        "roofline": {"bound": "hbm", "kernel": "integrate_kernel", "achieved": achieved, "peak": peak * world, "unit": "GB/s",
                     "frac": achieved / (peak * world), "traffic": None,
                     "traffic_source": "not captured at N > 1 (ncu runs on one GPU; see the N = 1 line)",
                     "peak_source": peak_src + f" x {world} GPUs",
                     "kernel_ms": k_max, "algorithmic_bytes": 8.0 * main["U_stored"]},
        "clocks": main["clocks"],
    }
    if strong is not None:
        s1 = strong["single"]
        ks = strong["stage_ms"]["integrate_kernel_slowest_rank"]
        line["strong"] = {
            "scaling": "strong", "workload": workload_label(sd) if workload_label else f"{sd}^3",
            "n_gpus": world, "ms_per_step": strong["ms_per_step"], "value": strong["U_owned"] / (strong["ms_per_step"] * 1e-3), "unit": UNIT,
            "single_gpu_ms_per_step": s1["ms_per_step"], "frame_speedup_vs_1gpu": s1["ms_per_step"] / strong["ms_per_step"],
            "integrate_kernel_ms_slowest_rank": ks, "single_gpu_integrate_kernel_ms": s1["integrate_kernel_ms_last_frame"],
            "integrate_kernel_speedup_vs_1gpu": s1["integrate_kernel_ms_last_frame"] / ks if ks > 0 else None,
            "roofline_frac_aggregate": 8.0 * strong["U_stored"] / (ks * 1e-3) / 1e9 / (peak * world) if ks > 0 else None,
            "stage_ms": strong["stage_ms"], "slab_bounds": strong["bounds"],
            "poses_equal_single_gpu": bool(s1["poses_equal"]),
        }
        parity["strong_final_pose_equal"] = bool(s1["poses_equal"])
    print(json.dumps(line))
    if not parity["final_pose_equal"] or not parity.get("strong_final_pose_equal", True):
        raise SystemExit("sharded run and single-GPU run disagree (see \"parity\" in the line above)")
