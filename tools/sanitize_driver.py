"""Small-shape run of every kernel on the hot path, for compute-sanitizer (tools/sanitize.sh).
64^3 volume, 320x240 frames: front end, per-iteration ICP kernel, integrate (3 poses, counting variant too),
raycast + model pyramid, extraction, render; then -- unless KFB_SAN_NO_PERSISTENT is set -- three frames through
the C++ facade, which runs the whole-schedule ICP kernel (raise its poll bound with KFB_ICP_TIMEOUT_NS: kernels are
10-100x slower under the sanitizer)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import slam_kinectfusion_b200 as kfb  # noqa: E402
from slam_kinectfusion_b200 import synth  # noqa: E402


def main():
    w, h, dims = 320, 240, 64
    s = w / 640.0
    K = kfb.Intrinsics(width=w, height=h, fx=525.0 * s, fy=525.0 * s, cx=(319.5 + 0.5) * s - 0.5, cy=(239.5 + 0.5) * s - 0.5)
    P = kfb.default_params(dims)
    volpose = np.array(kfb.default_host_params(dims).volu_pose, np.float32)
    ctx = kfb.Context(K, P)

    def v2c(pose):
        Pm = np.vstack([np.asarray(pose, np.float64).reshape(3, 4), [0, 0, 0, 1]])
        V = np.vstack([volpose.astype(np.float64).reshape(3, 4), [0, 0, 0, 1]])
        return (np.linalg.inv(Pm) @ V)[:3].astype(np.float32).reshape(12)

    def c2v(pose):
        Pm = np.vstack([np.asarray(pose, np.float64).reshape(3, 4), [0, 0, 0, 1]])
        V = np.vstack([volpose.astype(np.float64).reshape(3, 4), [0, 0, 0, 1]])
        M = (np.linalg.inv(V) @ Pm)
        return M[:3].astype(np.float32).reshape(12), np.ascontiguousarray(M[:3, :3].T, np.float32).reshape(9)

    for k in (0, 6, 12):
        pose = synth.trajectory_pose(k)
        d = synth.render_depth_mm(pose, w, h, K.fx, K.fy, K.cx, K.cy)
        ctx.upload_depth_mm(d)
        ctx.frontend()
        n = ctx.integrate(v2c(pose), count=(k == 6))
        m, r = c2v(pose)
        ctx.raycast(m, r)
        ctx.model_pyramid()
    ctx.swap_frames()
    ctx.upload_depth_mm(synth.render_depth_mm(synth.trajectory_pose(13), w, h, K.fx, K.fy, K.cx, K.cy))
    ctx.frontend()
    ident = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], np.float32)
    for l in (2, 1, 0):
        sums = ctx.icp_accumulate(l, ident)
    pts = ctx.extract_points(volpose)
    img = ctx.render_phong(np.zeros(3, np.float32))
    img2 = ctx.render_normals()
    vol = ctx.download_volume()
    ctx.upload_volume(vol)
    ctx.synchronize()
    print("direct kernels ok: updated", n, "points", len(pts), "sum0 %.6g" % sums[0], "render", int(img.any()), int(img2.any()))
    ctx.close()
    if os.environ.get("KFB_SAN_NO_PERSISTENT"):
        return
    Kh = kfb.Intrinsics(width=w, height=h, fx=K.fx, fy=K.fy, cx=K.cx, cy=K.cy)
    kf = kfb.KinectFusion(Kh, kfb.default_host_params(dims))
    for k in range(3):
        d = synth.render_depth_mm(synth.trajectory_pose(k), w, h, K.fx, K.fy, K.cx, K.cy)
        rc = kf.pipeline(d)
        assert rc == 0, rc
    print("facade ok: 3 frames, icp fallbacks", kf.context().icp_fallback_count(), "pose", kf.pose()[[3, 7, 11]])
    kf.close()


if __name__ == "__main__":
    main()
