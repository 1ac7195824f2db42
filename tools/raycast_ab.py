"""Raycast A/B inside the pipelined frame sequence: linear voxel layout against the brick-major copy
(KFB_RAYCAST_BLOCKED=1, north_star's "L2-friendly voxel layout"); raycast kernel duration from events 58/59,
results compared bit for bit.
    python tools/raycast_ab.py [dims] [frames]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import slam_kinectfusion_b200 as kfb  # noqa: E402
from slam_kinectfusion_b200 import synth  # noqa: E402


def main():
    dims = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    frames = synth.sequence(n, K)
    dev = torch.stack([torch.from_numpy(d) for _, d in frames]).cuda()
    kf = kfb.KinectFusion(K, kfb.default_host_params(dims))
    ctx = kf.context()
    ctx.set_profiling(True)
    out = {}
    for name, env in (("linear", None), ("brick-major 8^3", "1"), ("linear", None), ("brick-major 8^3", "1")):
        os.environ.pop("KFB_RAYCAST_BLOCKED", None)
        if env:
            os.environ["KFB_RAYCAST_BLOCKED"] = env
        kf.reset()
        ms = []
        for i in range(n):
            assert kf.pipeline_ptr(dev[i].data_ptr(), K.width, K.height) == 0
            if i >= 8 and i % 4 == 3:
                ctx.synchronize()
                ms.append(ctx.event_elapsed_ms(58, 59))
        ctx.synchronize()
        v, nm = ctx.download_maps(1, 0)
        out.setdefault(name, (kf.pose().copy(), v.copy(), nm.copy()))
        same = all(np.array_equal(a.view(np.int32), b.view(np.int32)) for a, b in zip(out[name][1:], (v, nm)))
        print(f"{name:18s} raycast kernel {np.mean(ms) * 1e3:7.1f} us  (min {np.min(ms) * 1e3:.1f})  repeatable {same}", flush=True)
    a, b = out["linear"], out["brick-major 8^3"]
    print("poses equal:", np.array_equal(a[0], b[0]), " model maps bit-identical:",
          np.array_equal(a[1].view(np.int32), b[1].view(np.int32)) and np.array_equal(a[2].view(np.int32), b[2].view(np.int32)))


if __name__ == "__main__":
    main()
