import sys, time
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import slam_kinectfusion_b200 as kfb
from slam_kinectfusion_b200 import synth
K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
ctx = kfb.Context(K, kfb.default_params(128))
f0 = synth.render_depth_mm(synth.trajectory_pose(0)); f1 = synth.render_depth_mm(synth.trajectory_pose(1))
ctx.upload_depth_mm(f0); ctx.frontend(); ctx.swap_frames(); ctx.upload_depth_mm(f1); ctx.frontend(); ctx.synchronize()
I = np.zeros(12, np.float32); I[0] = I[5] = I[10] = 1
iters = [4, 5, 10]
for trial in range(4):
    ts = []
    t0 = time.perf_counter(); ctx.icp_begin(iters); tb = time.perf_counter() - t0
    for k in range(19):
        t1 = time.perf_counter(); ctx.icp_step(I); ts.append((time.perf_counter() - t1) * 1e6)
    t2 = time.perf_counter(); ctx.icp_end(); te = time.perf_counter() - t2
    print("begin us", round(tb * 1e6, 1), "end us", round(te * 1e6, 1), "steps us", [round(t, 1) for t in ts], "total ms", round((time.perf_counter() - t0) * 1e3, 3))
for l in (2, 1, 0):
    t0 = time.perf_counter()
    for _ in range(50): ctx.icp_accumulate(l, I)
    print("direct level", l, (time.perf_counter() - t0) / 50 * 1e6, "us")


from slam_kinectfusion_b200 import host as H
for trial in range(3):
    us = H.icp_probe(ctx)
    print("C++ step us", [round(float(u), 1) for u in us], "total", round(float(us.sum()), 1))

# per-iteration stamps of the last schedule (CTA 0)
R = ctx.debug_icp_ring().astype(np.int64)[:19]
print("accumulate ns", (R[:, 1] - R[:, 0]).tolist())
print("block sums + exchange + final sums ns", (R[:, 2] - R[:, 1]).tolist())
print("post + next pose ns", (R[:-1, 3] - R[:-1, 2]).tolist())
print("period ns", np.diff(R[:, 0]).tolist())
