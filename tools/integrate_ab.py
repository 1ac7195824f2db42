"""Timing A/B of the integrate variants inside the pipelined frame sequence (profiling events: sweep kernels
60/61, whole call 56/57), and a bit-for-bit comparison of the resulting volumes.  Variants are switched through
the library's environment switches, which are read at every launch.
    python tools/integrate_ab.py [dims] [frames]"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import slam_kinectfusion_b200 as kfb  # noqa: E402
from slam_kinectfusion_b200 import synth  # noqa: E402

# NOTE: the GPU's clocks sag over the first seconds of sustained load, so variants that run later in the process look
# slower (the same variant measured 104 us as the first and 117 us as the fourth of a run): repeat the variants of
# interest at both ends of the list, or compare runs of the whole tool.
VARIANTS = {
    "default": {},
    "no L2 prefetch": {"KFB_GEN_NOPREFETCH": "1"},
    "states kernel of its own": {"KFB_INTEGRATE_SPLITSTATES": "1"},
    "serial (one stream)": {"KFB_INTEGRATE_SERIAL": "1"},
    "chunks of 8 planes": {"KFB_PLAN_ZCHUNK": "8"},
    "general kernel 80 regs": {"KFB_GEN_MINB": "6"},
    "default (again)": {},
    "no L2 prefetch (again)": {"KFB_GEN_NOPREFETCH": "1"},
}
SWITCHES = ("KFB_GEN_NOPREFETCH", "KFB_INTEGRATE_SERIAL", "KFB_INTEGRATE_PERSISTENT", "KFB_GEN_MINB", "KFB_PLAN_ZCHUNK", "KFB_GEN_WARPS", "KFB_INTEGRATE_SPLITSTATES")


def main():
    dims = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    only = sys.argv[3] if len(sys.argv) > 3 else None      # run a single variant (profiler runs)
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    frames = synth.sequence(n, K)
    dev = torch.stack([torch.from_numpy(d) for _, d in frames]).cuda()
    hp = kfb.default_host_params(dims)
    kf = kfb.KinectFusion(K, hp)
    ctx = kf.context()
    ctx.set_profiling(True)
    digests = {}
    for name, env in VARIANTS.items():
        if only and only not in name:
            continue
        for k in SWITCHES:
            os.environ.pop(k, None)
        os.environ.update(env)
        kf.reset()
        k_ms, c_ms, f_ms = [], [], []
        for i in range(n):
            if i == 8:
                ctx.synchronize()
                ctx.event_record(0)
            assert kf.pipeline_ptr(dev[i].data_ptr(), K.width, K.height) == 0
            if i >= 8 and i % 3 == 2:
                ctx.synchronize()
                k_ms.append(ctx.event_elapsed_ms(60, 61))
                c_ms.append(ctx.event_elapsed_ms(56, 57))
        ctx.event_record(1)
        ctx.synchronize()
        vol = ctx.download_volume()
        digests[name] = hashlib.sha1(vol.tobytes()).hexdigest()
        # counts at the last pose
        ctx.upload_depth_mm_ptr(dev[n - 1].data_ptr(), K.width, K.height)
        ctx.frontend()
        P = np.vstack([kf.pose().astype(np.float64).reshape(3, 4), [0, 0, 0, 1]])
        V = np.vstack([np.array(hp.volu_pose, np.float64).reshape(3, 4), [0, 0, 0, 1]])
        U = ctx.integrate((np.linalg.inv(P) @ V)[:3].astype(np.float32).reshape(12), count=True)
        cnt = ctx.integrate_counts()
        print(f"{name:24s} sweep median {np.median(k_ms) * 1e3:6.1f} (min {np.min(k_ms) * 1e3:6.1f}) us  call median {np.median(c_ms) * 1e3:6.1f} us  "
              f"U {U}  8U/t {8 * U / np.median(k_ms) / 1e6:7.1f} GB/s  {cnt}  sha1 {digests[name][:12]}", flush=True)
    assert len(set(digests.values())) == 1, digests
    print("volumes identical across variants")


if __name__ == "__main__":
    main()
