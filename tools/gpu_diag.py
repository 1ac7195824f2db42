"""Scratch diagnostics run on the GPU box (prints stage timings and parity stats)."""
import sys, time, json
import numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import slam_kinectfusion_b200 as kfb
from oracle import kfo

def main():
    dims = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    Ko = kfo.intr(); Kb = kfb.Intrinsics(**kfb.SENSORS["kinect1"])
    ctx = kfb.Context(Kb, kfb.default_params(dims))
    volpose = np.array(kfo.default_params(dims).volu_pose, np.float32)
    frames = [kfo.render_depth_mm(kfo.trajectory_pose(k), Ko) for k in range(6)]
    # stage timings
    def timed(fn, reps=20):
        fn(); ctx.synchronize()
        ts = []
        for _ in range(reps):
            ctx.event_record(0); fn(); ctx.event_record(1)
            ts.append(ctx.event_elapsed_ms(0, 1))
        return float(np.median(ts)), float(np.min(ts))
    ctx.upload_depth_mm(frames[0])
    print("frontend ms", timed(ctx.frontend))
    cam = kfo.trajectory_pose(40)
    ctx.upload_depth_mm(kfo.render_depth_mm(cam, Ko)); ctx.frontend()
    v2c = kfo.pose_mul(kfo.pose_inv(cam), volpose)
    U = ctx.integrate(v2c, count=True)
    print("U", U)
    t = timed(lambda: ctx.integrate(v2c))
    print("integrate ms", t, "GB/s (8U/t)", 8 * U / (t[0] * 1e-3) / 1e9)
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), cam)
    rinv = kfo.rot_inv(c2v)
    print("raycast ms", timed(lambda: ctx.raycast(c2v, rinv)))
    print("pyramid ms", timed(ctx.model_pyramid))
    ctx.swap_frames(); ctx.upload_depth_mm(frames[1]); ctx.frontend()
    for l in (2, 1, 0):
        t0 = time.perf_counter()
        for _ in range(50): ctx.icp_accumulate(l, kfo.identity())
        print("icp level", l, "us/iter (host wall)", (time.perf_counter() - t0) / 50 * 1e6)
    # whole pipeline via facade
    kf = kfb.KinectFusion(Kb, kfb.default_host_params(dims))
    for k in range(3): kf.pipeline(frames[k])
    t0 = time.perf_counter()
    for k in range(3, 6): kf.pipeline(frames[k])
    kf.context().synchronize()
    print("pipeline ms/frame (wall)", (time.perf_counter() - t0) / 3 * 1e3)

main()
