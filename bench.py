#!/usr/bin/env python
"""bench.py -- the KinectFusion hot path on B200, measured per the driver contract.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one full frame of the hot path over one synthetic 640x480 depth image:
front end (pyrDown, bilateral, vertex/normal maps) + 3-level ICP (10/5/4 iterations, host 6x6
solve) + TSDF integration + raycast + model-map pyramid.

  N = 1   workload = BASELINE.json configs[1]: 640x480 depth, 512^3 TSDF over 3 m, 300-frame looped
          synthetic trajectory (analytic box+sphere room).  `ms_per_step` IS the headline
          "frame device-ms at 640x480/512^3".
  N > 1   the volume is the only part of the path that shards: z-slab sharded
          volume with ~512^3 voxels per GPU (640^3 / 812^3 / 1024^3 at N = 2 / 4 / 8), every rank
          integrating its slab of the same frame with no data-path collective => "weak".

`value` = TSDF voxel-updates/s sustained over whole frames = (voxels passing the reference's update
predicate per frame, summed over ranks) / (device time per frame, max over ranks).  Both arms see
the same frames, so the ratio of the two arms' values is the frame-rate ratio.  The kernel-only
integration rate is in `roofline` (8 B per updated voxel against the measured HBM copy bandwidth).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "TSDF voxel-updates/s sustained over full frames (frame device-ms at 640x480/512^3 = ms_per_step)"
UNIT = "voxel-updates/s"
WEAK_DIMS = {1: 512, 2: 640, 4: 812, 8: 1024}


TRAFFIC_FILE = "r02_integrate_traffic.json"


def ncu_traffic():
    """dram bytes per launch of the integrate kernel.  NOT measured in this run (ncu cannot run inside a timed
    bench): read from the committed `ncu --set full` capture of this same command under profiles/."""
    for name in (TRAFFIC_FILE, "r01_integrate_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["dram_bytes_per_launch"]), f"committed ncu capture profiles/{name} (not measured in this run)"
        except Exception:
            continue
    return None, "no capture committed"


def workload_label(dims):
    """One string for both arms (the driver compares them): the job, not how an arm runs it."""
    return (f"640x480 depth, {dims}^3 TSDF over 3 m, ICP 10/5/4, 300-frame looped synthetic box+sphere room trajectory"
            + (" (BASELINE configs[1])" if dims == 512 else ""))


def measured_peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def oracle_frames_per_s(dims, frames, threads):
    """CPU baseline: the oracle's whole pipeline (oracle/kf_oracle.c, OpenMP) on a bounded sample."""
    from oracle import kfo
    kfo.build()
    os.environ["OMP_NUM_THREADS"] = str(threads)
    K = kfo.intr()
    kf = kfo.Kinfu(K, kfo.default_params(dims))
    t_frames, U = [], []
    for i, (pose, d) in enumerate(frames):
        t0 = time.perf_counter()
        rc = kf.pipeline(d)
        dt = time.perf_counter() - t0
        assert rc == 0
        if i >= 1:                 # frame 1 is the bootstrap (no ICP / raycast), like the reference's timer
            t_frames.append(dt)
            U.append(kf.last_updated)
    return float(np.mean(U)) / float(np.mean(t_frames)), float(np.mean(t_frames)), float(np.mean(U))


def run_reference(args):
    """`--impl reference`: the reference's implementation of the path on this box, none of this repo's kernels
    on it.  The reference is CUDA-only, so when its kernels were compiled here (oracle/_ref/libkf_ref.so: the
    three .cu files of the reference, unmodified, for sm_100a) and a GPU is present, this arm runs the
    reference's own frame loop over ITS kernels on the same B200 (oracle/ref_harness.cu: its 8-byte voxel, its
    per-iteration cudaMalloc/memcpy in rigidICP, its cudaDeviceSynchronize; the two un-vendored OpenCV-CUDA
    calls are stood in for by plain kernels in the harness) -- kind "reference".  Otherwise it is the
    scalar/OpenMP oracle port on the host cores -- kind "port".  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from slam_kinectfusion_b200 import synth
    import slam_kinectfusion_b200 as kfb
    # the same job as the other arm at this N (the weak-scaled volume); the reference cannot shard, so it runs
    # the whole volume on one GPU
    dims = WEAK_DIMS.get(args.gpus, 512) if args.dims is None else args.dims
    cores = os.cpu_count() or 1
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    W, S = args.warmup, args.steps
    t_start = time.perf_counter()
    from oracle import kref, kfo
    gpu = False
    if kref.available():
        try:
            import torch
            gpu = torch.cuda.is_available()
        except Exception:
            gpu = False
    workload = workload_label(dims)
    if gpu:
        import torch
        S = min(S, 100)
        n = 1 + W + S
        frames = synth.sequence(n, K)
        pin = torch.empty((n, K.height, K.width), dtype=torch.float32).pin_memory()
        for i, (_, d) in enumerate(frames):
            pin[i].copy_(torch.from_numpy(d))
        volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
        rk = kref.RefKinfu(K, dims, volpose)
        for i in range(n):
            if i == 1 + W:
                kref.lib().ref_device_sync()
                kref.event_tic()
                t0 = time.perf_counter()
            if rk.pipeline_ptr(pin[i].data_ptr()) != 0:
                raise SystemExit(f"reference pipeline lost tracking at frame {i}")
        ms = kref.event_toc_ms()
        wall = time.perf_counter() - t0
        ms_per_frame = max(ms, wall * 1e3) / S
        # updated voxels per frame: a property of the workload (frames + poses); counted by the oracle on a sample
        Ko = kfo.intr()
        vd = kfo.volume_desc(dims)
        U = []
        for i in ((1 + W, n - 1) if dims <= 512 else (1 + W,)):
            vol = kfo.new_volume(vd)
            dm = kfo.frontend(frames[i][1], Ko, levels=1)[0][0]
            U.append(kfo.integrate(vol, vd, kfo.pose_mul(kfo.pose_inv(frames[i][0]), volpose), dm, Ko))
            del vol
        U_mean = float(np.mean(U))
        value = U_mean / (ms_per_frame * 1e-3)
        kind, ncores = "reference", 1
        sample = (f"{S} frames after {W} warm-up frames: the reference's own CUDA kernels (oracle/_ref, sm_100a rebuild) "
                  f"driven by its frame loop on this GPU, host frames in pinned memory, 1 host thread; {ms_per_frame:.3f} ms/frame")
        steps = S
    else:
        n = min(1 + W + S, 1 + 2 + 6)         # bounded sample: ~0.35 s per 512^3 frame on 16 cores
        frames = synth.sequence(n, K)
        value, sec_per_frame, U_mean = oracle_frames_per_s(dims, frames, cores)
        ms_per_frame = sec_per_frame * 1e3
        kind, ncores = "port", cores
        sample = f"{n - 1} frames after the bootstrap frame, whole pipeline (oracle/kf_oracle.c), OpenMP over {cores} threads"
        steps, W = n - 1, 0
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": W, "ms_per_step": ms_per_frame, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 -> int16 tsdf", "data": "synthetic",
        "config": {"workload": workload, "l2": "inputs larger than L2 (volume swept every frame)",
                   "updated_voxels_per_frame": U_mean},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": kind, "sample": sample},
        # the contract fixes the two byte counts of this arm at 0; what the reference loop really moves per frame
        # is stated next to them
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "actual_h2d_bytes_per_step": (K.width * K.height * 4) if kind == "reference" else 0,
                "actual_d2h_bytes_per_step": (19 * 27 * 4) if kind == "reference" else 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t_start,
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import slam_kinectfusion_b200 as kfb
    from slam_kinectfusion_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    dims = args.dims if args.dims else WEAK_DIMS.get(world, 512)
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    W, S = args.warmup, args.steps
    n_frames = 1 + W + S                       # frame 0 bootstraps the volume (no ICP / raycast)
    frames = synth.sequence(n_frames, K)
    w, h = K.width, K.height
    frame_bytes = w * h * 4

    # inputs: pinned host copies (e2e) and device-resident copies (device-timed value)
    host_pin = torch.empty((n_frames, h, w), dtype=torch.float32).pin_memory()
    for i, (_, d) in enumerate(frames):
        host_pin[i].copy_(torch.from_numpy(d))
    dev_frames = host_pin.to(f"cuda:{local}", non_blocking=False)
    torch.cuda.synchronize()

    hp = kfb.default_host_params(dims)
    hp.device = local
    if world > 1:
        from slam_kinectfusion_b200 import sharded
        try:
            return sharded.run_bench(args, dist, rank, world, local, dims, K, frames, host_pin, dev_frames,
                                     METRIC, UNIT, measured_peak_hbm, ClockSampler, workload_label)
        finally:
            dist.barrier()
            dist.destroy_process_group()

    kf = kfb.KinectFusion(K, hp)
    ctx = kf.context()

    def run(seq_ptrs, timed_from):
        """Feed frames; returns (device ms over the timed part, wall s over the timed part, launches)."""
        kf.reset()
        l0 = t0 = None
        for i, ptr in enumerate(seq_ptrs):
            if i == timed_from:
                ctx.synchronize()
                l0 = ctx.launch_count()
                ctx.event_record(0)
                t0 = time.perf_counter()
            rc = kf.pipeline_ptr(ptr, w, h)
            if rc != 0:
                raise SystemExit(f"pipeline rc={rc} at frame {i} ({'tracking failure' if rc == 1 else kf.last_error()})")
            _ = kf.pose()                      # the frame's result on the host (ICP sums already crossed PCIe)
        ctx.event_record(1)
        ctx.synchronize()
        wall = time.perf_counter() - t0
        return ctx.event_elapsed_ms(0, 1), wall, ctx.launch_count() - l0

    dptr = [dev_frames[i].data_ptr() for i in range(n_frames)]
    hptr = [host_pin[i].data_ptr() for i in range(n_frames)]

    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, _, launches = run(dptr, 1 + W)
    clocks = sampler.stop()
    ms_per_frame = dev_ms / S
    e2e_ms, e2e_wall, _ = run(hptr, 1 + W)
    e2e_ms_per_frame = max(e2e_ms, e2e_wall * 1e3) / S
    poses = kf.poses()

    # ---- outside the timed regions: (1) the stage kernels' own durations, measured IN SITU: the same pipelined
    # sequence once more with the library's profiling events on (integrate kernel 60/61, whole integrate call
    # 56/57, raycast 58/59); every 4th frame is followed by a synchronize so that its events can be read -- the
    # kernels of that frame ran in their normal context (cold volume, tables and states just written)
    ctx.set_profiling(True)
    kf.reset()
    k_ms, call_ms, rc_ms, icp_ms = [], [], [], []
    for i, ptr in enumerate(dptr):
        rc = kf.pipeline_ptr(ptr, w, h)
        if rc != 0:
            raise SystemExit(f"pipeline rc={rc} at frame {i} of the profiled pass ({'tracking failure' if rc == 1 else kf.last_error()})")
        if i > W and i % 4 == 0:
            ctx.synchronize()
            k_ms.append(ctx.event_elapsed_ms(60, 61))
            call_ms.append(ctx.event_elapsed_ms(56, 57))
            rc_ms.append(ctx.event_elapsed_ms(58, 59))
            try:
                icp_ms.append(ctx.event_elapsed_ms(54, 55))
            except Exception:  # KFB_ICP_DIRECT: no whole-schedule kernel, the events were never recorded
                pass
    ctx.set_profiling(False)
    # (2) updated-voxel counts (the counting variant of the kernel, on sampled frames at their tracked poses)
    volpose = np.array(hp.volu_pose, np.float32).reshape(3, 4)

    def vol2cam(p12):
        P = np.vstack([np.asarray(p12, np.float64).reshape(3, 4), [0, 0, 0, 1]])
        V = np.vstack([volpose.astype(np.float64), [0, 0, 0, 1]])
        return (np.linalg.inv(P) @ V)[:3].astype(np.float32).reshape(12)

    U, moved = [], []
    for i in range(1 + W, n_frames, max(1, S // 16)):
        ctx.upload_depth_mm_ptr(dptr[i], w, h)
        ctx.frontend()
        U.append(ctx.integrate(vol2cam(poses[i]), count=True))
        c = ctx.integrate_counts()
        moved.append(16.0 * (c["quads_loaded"] + c["quads_stored"]))   # the sweep's own count of the bytes it loads and stores
    U_mean = float(np.mean(U))
    k_ms_mean = float(np.mean(k_ms))
    peak, peak_src = measured_peak_hbm()
    achieved = 8.0 * U_mean / (k_ms_mean * 1e-3) / 1e9
    swept = dims * dims * (dims - 1)

    # ---- steady state (outside the timed region): the 300-frame configuration spends most of its frames with the
    # weights of everything in view saturated at 64; free-space voxels then keep their value and their store is
    # dropped, so 8 B per updated voxel overstates what the sweep moves.  Age the volume by 72 more frames of the same
    # trajectory, then measure the sweep kernels' time and their own byte counts.
    steady = None
    try:
        extra = synth.sequence(72 + 12, K, start=n_frames)
        edev = torch.stack([torch.from_numpy(d) for _, d in extra]).to(f"cuda:{local}")
        sk = []
        for i in range(len(extra)):
            if kf.pipeline_ptr(edev[i].data_ptr(), w, h) != 0:
                raise RuntimeError(f"tracking failure at aged frame {i}")
            if i >= 72 and i % 4 == 3:
                ctx.synchronize()
                sk.append(ctx.event_elapsed_ms(60, 61))
        ctx.synchronize()
        sp = kf.pose()
        ctx.upload_depth_mm_ptr(edev[-1].data_ptr(), w, h)
        ctx.frontend()
        sU = ctx.integrate(vol2cam(sp), count=True)
        sc = ctx.integrate_counts()
        sbytes = 16.0 * (sc["quads_loaded"] + sc["quads_stored"])
        sms = float(np.mean(sk))
        steady = {"frames_integrated_before": n_frames + 72, "kernel_ms": sms, "updated_voxels": int(sU), "bytes_moved": sbytes,
                  "achieved_8B_per_update": 8.0 * sU / (sms * 1e-3) / 1e9, "achieved_bytes_moved": sbytes / (sms * 1e-3) / 1e9,
                  "frac_8B_per_update": 8.0 * sU / (sms * 1e-3) / 1e9 / peak, "frac_bytes_moved": sbytes / (sms * 1e-3) / 1e9 / peak,
                  "unit": "GB/s"}
        del edev
    except Exception as e:  # noqa: BLE001 - a side measurement
        steady = {"error": str(e)}

    # ---- dense-update micro-config: wide camera (fx = fy = 80) facing a wall at 4 m behind the
    # volume => every swept voxel is updated with tsdf = 1; isolates the kernel's HBM streaming rate
    dense = None
    try:
        kf.close()
        Kw = kfb.Intrinsics(width=w, height=h, fx=80.0, fy=80.0, cx=319.5, cy=239.5)
        cw = kfb.Context(Kw, kfb.default_params(dims), device=local)
        cw.upload_depth_mm(np.full((h, w), 4000.0, np.float32))
        cw.frontend()
        cw.set_profiling(True)
        Ud = cw.integrate(volpose.reshape(12), count=True)
        dk = []
        for _ in range(12):
            cw.integrate(volpose.reshape(12))
            dk.append(cw.event_elapsed_ms(60, 61))
        dms = float(np.median(dk[2:]))
        dense = {"updated_voxels": int(Ud), "swept_voxels": swept, "kernel_ms": dms,
                 "achieved": 8.0 * Ud / (dms * 1e-3) / 1e9, "unit": "GB/s", "frac": 8.0 * Ud / (dms * 1e-3) / 1e9 / peak}
        cw.close()
    except Exception as e:  # noqa: BLE001 - the micro-config is a side measurement
        dense = {"error": str(e)}

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 / N = 1 only
    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        # ~10-15 s of host work: ~0.33 s per frame at 512^3 on 16 cores, cubic in the volume side
        nb = 1 + max(3, min(40, int(40 * (512.0 / dims) ** 3)))
        nb = min(nb, len(frames))
        cval, csec, cU = oracle_frames_per_s(dims, frames[:nb], cores)
        cpu = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nb - 1} frames after the bootstrap frame of the same sequence, whole pipeline "
                         f"(oracle/kf_oracle.c, OpenMP {cores} threads), {csec * 1e3:.0f} ms/frame"}

    line = {
        "metric": METRIC, "value": U_mean / (ms_per_frame * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": S, "warmup": W,
        "ms_per_step": ms_per_frame, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 -> int16 tsdf", "data": "synthetic",
        "config": {"workload": workload_label(dims),
                   "l2": f"inputs larger than L2: the {dims ** 3 * 4 >> 20} MiB volume is swept by integrate and raycast every frame",
                   "frames_timed": S, "updated_voxels_per_frame": U_mean, "swept_voxels_per_frame": swept},
        "frame_device_ms": ms_per_frame,
        "e2e": {"value": U_mean / (e2e_ms_per_frame * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_per_frame,
                "device_ms_per_step": e2e_ms / S, "wall_ms_per_step": e2e_wall * 1e3 / S,
                "h2d_bytes_per_step": frame_bytes, "d2h_bytes_per_step": 19 * 27 * 16,
                "api": "kf::kinectfusion::pipeline(depth_mm) via libkfusion_b200.so, pinned host frames"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "integrate sweep", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1], "peak_source": peak_src,
                     "kernel_ms": k_ms_mean, "kernel_ms_how": "CUDA events around the launch, in situ in the pipelined sequence",
                     "integrate_call_ms": float(np.mean(call_ms)), "raycast_kernel_ms": float(np.mean(rc_ms)), "icp_kernel_ms": float(np.mean(icp_ms)) if icp_ms else None,
                     "algorithmic_bytes": 8.0 * U_mean, "bytes_moved_counted": float(np.mean(moved)),
                     "kernels": "integrate_general_kernel (its first blocks walk the running sums; second stream) || integrate_stream_kernel, events around both",
                     "steady_state": steady,
                     "dense_model_gbs": 8.0 * swept / (k_ms_mean * 1e-3) / 1e9, "dense_microconfig": dense},
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strong-dims", type=int, default=2048,
                    help="N > 1: also run this volume (BASELINE configs[3]/[4]) on the N GPUs and on one GPU; 0 = skip")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
