"""Scratch diagnostics on the GPU box: per-frame ICP wall time and misprediction count through the C++ facade."""
import sys
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import slam_kinectfusion_b200 as kfb
from slam_kinectfusion_b200 import synth
K = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
kf = kfb.KinectFusion(K, kfb.default_host_params(512))
frames = synth.sequence(40, K)
us = []
for i, (_, d) in enumerate(frames):
    assert kf.pipeline(d) == 0
    if i >= 5:
        us.append(kf.last_icp_us())
ctx = kf.context()
print("icp us per frame: median %.1f min %.1f max %.1f" % (np.median(us), np.min(us), np.max(us)))
print("mispredicted iterations since creation:", ctx.icp_mispredict_count(), "of", 19 * 39)

# serialized frames (sync between frames): ICP wall time is pure, frame wall = sum of the stages
import time
kf.reset()
fr, ic = [], []
for i, (_, d) in enumerate(frames):
    ctx.synchronize()
    t0 = time.perf_counter()
    assert kf.pipeline(d) == 0
    ctx.synchronize()
    if i >= 5:
        fr.append((time.perf_counter() - t0) * 1e6)
        ic.append(kf.last_icp_us())
print("serialized: frame us median %.1f, icp us median %.1f, rest %.1f" % (np.median(fr), np.median(ic), np.median(fr) - np.median(ic)))

R = ctx.debug_icp_ring().astype(np.int64)[:19]
print("per-iteration ns: accumulate", (R[:, 1] - R[:, 0]).tolist())
print("block sums + exchange + final sums", (R[:, 2] - R[:, 1]).tolist())
print("post + next pose", (R[:-1, 3] - R[:-1, 2]).tolist())
print("pose ready -> next entry", (R[1:, 0] - R[:-1, 3]).tolist())
print("period", np.diff(R[:, 0]).tolist())

# device-side stage times of pipelined frames (events recorded by the library when profiling is on)
kf.reset()
ctx.set_profiling(True)
rows = []
for i, (_, d) in enumerate(frames):
    assert kf.pipeline(d) == 0
    if i >= 5:
        ctx.synchronize()      # events of this frame are complete (serialises the frames: stage times only)
        rows.append((ctx.event_elapsed_ms(56, 57), ctx.event_elapsed_ms(60, 61), ctx.event_elapsed_ms(57, 58), ctx.event_elapsed_ms(58, 59)))
ctx.set_profiling(False)
R = np.array(rows) * 1e3
print("us: integrate call %.1f (kernel %.1f), gap to raycast %.1f, raycast %.1f" % tuple(np.median(R, axis=0)))

# the same, pipelined (a synchronize only after every 4th frame, whose events are then read)
kf.reset()
ctx.set_profiling(True)
rows = []
for i, (_, d) in enumerate(frames):
    assert kf.pipeline(d) == 0
    if i >= 5 and i % 4 == 0:
        ctx.synchronize()
        rows.append((ctx.event_elapsed_ms(54, 55), ctx.event_elapsed_ms(55, 56), ctx.event_elapsed_ms(56, 57), ctx.event_elapsed_ms(57, 58),
                     ctx.event_elapsed_ms(58, 59), ctx.event_elapsed_ms(54, 59)))
ctx.set_profiling(False)
R = np.array(rows) * 1e3
print("pipelined us: icp kernel %.1f, gap %.1f, integrate call %.1f, gap %.1f, raycast %.1f; icp start -> raycast end %.1f" % tuple(np.median(R, axis=0)))
