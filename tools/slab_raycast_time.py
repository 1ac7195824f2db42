"""Raycast of single z-slabs of a large volume on ONE GPU (slab contexts, no peers): kernel time per slab.
    python tools/slab_raycast_time.py [dims]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam_kinectfusion_b200 as kfb  # noqa: E402
from slam_kinectfusion_b200 import synth  # noqa: E402


def mat(p12):
    return np.vstack([np.asarray(p12, np.float64).reshape(3, 4), [0, 0, 0, 1]])


def main():
    dims = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    V = mat(kfb.default_host_params(dims).volu_pose)
    cams = [synth.trajectory_pose(k) for k in (0, 6, 12)]
    depths = [synth.render_depth_mm(c) for c in cams]
    cut = [0, 0.38, 0.5, 0.6, 0.67, 0.74, 0.81, 0.875, 1.0]
    for i in range(8):
        z0, z1 = int(cut[i] * dims), int(cut[i + 1] * dims)
        P = kfb.default_params(dims)
        P.slab_z_begin, P.slab_z_end = z0, z1
        ctx = kfb.Context(K, P)
        for c, d in zip(cams, depths):
            ctx.upload_depth_mm(d)
            ctx.frontend()
            ctx.integrate((np.linalg.inv(mat(c)) @ V)[:3].astype(np.float32).reshape(12))
        C2V = np.linalg.inv(V) @ mat(synth.trajectory_pose(14))
        c2v = C2V[:3].astype(np.float32).reshape(12)
        rinv = np.ascontiguousarray(C2V[:3, :3].T.astype(np.float32)).reshape(9)
        ctx.set_profiling(True)
        ts = []
        for _ in range(12):
            ctx.raycast(c2v, rinv)
            ctx.synchronize()
            ts.append(ctx.event_elapsed_ms(58, 59))
        gv, _ = ctx.download_maps(1, 0)
        print(f"slab [{z0:4d},{z1:4d}) of {dims}: pixels with an event here {(gv[..., 2] != 0).mean():.2f}  raycast {np.median(ts[2:]) * 1e3:7.1f} us", flush=True)
        ctx.close()


if __name__ == "__main__":
    main()
