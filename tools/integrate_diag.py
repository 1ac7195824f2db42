"""Timing decomposition of the integrate kernel (GPU box).  KFB_INTEGRATE_DIAG: 1 = general-path warps return,
2 = fast-path planes are skipped, 3 = both (per-warp setup only).  With the switch set results are WRONG; this tool
never checks them, it only reads the in-situ event timers."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam_kinectfusion_b200 as kfb
from oracle import kfo

def main():
    dims = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    chunks = [int(c) for c in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    Ko = kfo.intr(); Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    poses = [kfo.trajectory_pose(30 + k) for k in range(12)]
    frames = [kfo.render_depth_mm(p, Ko) for p in poses]
    for zc in chunks:
        for diag in ((0, 3, 1, 2) if len(chunks) == 1 else (0,)):
            os.environ.pop("KFB_INTEGRATE_DIAG", None)
            os.environ.pop("KFB_INTEGRATE_ZCHUNKS", None)
            if zc: os.environ["KFB_INTEGRATE_ZCHUNKS"] = str(zc)
            ctx = kfb.Context(Kb, kfb.default_params(dims))
            ctx.set_profiling(True)
            for k in range(8):   # build the scene with the real kernel
                ctx.upload_depth_mm(frames[k]); ctx.frontend()
                ctx.integrate(kfo.pose_mul(kfo.pose_inv(poses[k]), volpose))
            if diag: os.environ["KFB_INTEGRATE_DIAG"] = str(diag)
            tk, tc = [], []
            for k in range(8, 12):
                ctx.upload_depth_mm(frames[k]); ctx.frontend()
                ctx.integrate(kfo.pose_mul(kfo.pose_inv(poses[k]), volpose)); ctx.synchronize()
                tk.append(ctx.event_elapsed_ms(60, 61)); tc.append(ctx.event_elapsed_ms(56, 57))
            print("zchunks", zc or "default", "diag", diag, "kernel ms", np.round(tk, 4), "call ms", np.round(tc, 4), flush=True)
            del ctx
main()
