"""Generates tests/golden/oracle_golden.npz from the CPU oracle (oracle/kf_oracle.c).

The reference ships no golden vectors, so the fixtures are oracle outputs on
one small deterministic case; they pin the oracle against regressions and are what the GPU path
is compared with in tests/test_gpu_parity.py::test_golden_fixture.  Re-run:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kfo  # noqa: E402


def main():
    K = kfo.intr(160, 128, 131.25, 131.25, 79.5, 63.5)
    dims = 48
    pose0, pose1 = kfo.trajectory_pose(0), kfo.trajectory_pose(6)
    d0 = kfo.render_depth_mm(pose1, K)
    # a few holes and a far pixel so the invalid-depth paths are exercised
    d0[10:14, 20:26] = 0.0
    d0[60, 80] = 6000.0
    fe = kfo.frontend(d0, K)
    vd = kfo.volume_desc(dims)
    P = kfo.default_params(dims)
    volpose = np.array(P.volu_pose, np.float32)
    v2c = kfo.pose_mul(kfo.pose_inv(pose1), volpose)
    vol = kfo.new_volume(vd)
    # integrate the same view three times so weights/decay are exercised, plus a first view
    fe_first = kfo.frontend(kfo.render_depth_mm(pose0, K), K, levels=1)
    d_first = fe_first[0][0]
    kfo.integrate(vol, vd, kfo.pose_mul(kfo.pose_inv(pose0), volpose), d_first, K)
    pre = vol.copy()
    kfo.integrate(vol, vd, v2c, fe[0][0], K)
    c2v = kfo.pose_mul(kfo.pose_inv(volpose), pose1)
    v, n, steps = kfo.raycast(vol, vd, c2v, K)
    icp_pose = kfo.pose_apply_increment(kfo.identity(), np.array([0.002, -0.001, 0.0015, 0.003, -0.002, 0.001]))
    # ICP: current maps against the measured maps of the first view (the frame-2 case)
    s27, cnt = kfo.icp_accumulate(fe[0][1], fe[0][2], fe_first[0][1], fe_first[0][2], K, icp_pose)
    pts = kfo.extract_points(vol, vd, volpose)
    out = dict(intr=np.array([K.width, K.height, K.fx, K.fy, K.cx, K.cy], np.float64), dims=np.int32(dims),
               depth_mm=d0, vol2cam=v2c, cam2vol=c2v, volpose=volpose, volume_pre=pre, volume=vol,
               ray_v=v, ray_n=n, icp_pre_v=fe_first[0][1], icp_pre_n=fe_first[0][2], ray_steps=np.int64(steps), icp_pose=icp_pose, icp_sums=s27,
               icp_count=np.int64(cnt), points=pts)
    for l in range(3):
        out[f"depth_l{l}"], out[f"vmap_l{l}"], out[f"nmap_l{l}"] = fe[l]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", "icp count", cnt, "points", len(pts), "ray steps", steps)


if __name__ == "__main__":
    main()
