"""Soak run on the GPU box: the full 300-frame looped trajectory (BASELINE configs[1]) through the facade;
prints tracking error against the analytic ground truth and the ICP misprediction count."""
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam_kinectfusion_b200 as kfb
from slam_kinectfusion_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
kf = kfb.KinectFusion(K, kfb.default_host_params(512))
err_t, err_r = [], []
t0 = time.perf_counter()
for k in range(n):
    gt = synth.trajectory_pose(k)
    d = synth.render_depth_mm(gt)
    rc = kf.pipeline(d)
    assert rc == 0, f"tracking failure at frame {k}"
    p = kf.pose().reshape(3, 4).astype(np.float64)
    g = gt.reshape(3, 4).astype(np.float64)
    err_t.append(np.abs(p[:, 3] - g[:, 3]).max())
    dR = g[:, :3].T @ p[:, :3]
    err_r.append(0.5 * np.linalg.norm([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]]))
ctx = kf.context()
ctx.synchronize()
print(f"{n} frames ok; translation error vs ground truth: median {np.median(err_t) * 1e3:.2f} mm, max {np.max(err_t) * 1e3:.2f} mm; "
      f"rotation: median {np.median(err_r) * 1e3:.3f} mrad, max {np.max(err_r) * 1e3:.3f} mrad")
print("mispredicted ICP iterations:", ctx.icp_mispredict_count(), "of", 19 * (n - 1), "; schedules that fell back to ordinary launches:", ctx.icp_fallback_count())
pts = kf.extract_pointcloud()
print("point cloud:", len(pts), "points; bbox", pts.min(0).round(3), pts.max(0).round(3))
