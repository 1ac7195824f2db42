"""Headless version of the reference's main.cpp loop (main.cpp:64-116) over a dataset directory
(color/*.png, depth/*.png 16-bit mm, intr.txt): track + map every frame, then write poses.txt and the point cloud.

    python tools/run_dataset.py <dataset dir> [--dims 512] [--out out_dir]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import slam_kinectfusion_b200 as kfb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dataset")
    ap.add_argument("--dims", type=int, default=512)
    ap.add_argument("--out", default=".")
    a = ap.parse_args()
    cam = kfb.DatasetSensor(a.dataset)
    if cam.fx <= 0:
        raise SystemExit("intr.txt missing or malformed (expected fx 0 cx / 0 fy cy / 0 0 1)")
    K = kfb.Intrinsics(width=cam.width, height=cam.height, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy)
    kinfu = kfb.KinectFusion(K, kfb.default_host_params(a.dims))
    n = fails = 0
    for _, depth_mm in cam:
        fails += kinfu.pipeline(depth_mm) != 0  # "tracking fail!" resets, as in the reference
        n += 1
    os.makedirs(a.out, exist_ok=True)
    kinfu.save_poses(os.path.join(a.out, "poses.txt"))
    kinfu.extract_pointcloud()
    kinfu.save_pointcloud(os.path.join(a.out, "scene.ply"))
    print(f"{n} frames, {fails} tracking failures -> {a.out}/poses.txt, scene.ply")


if __name__ == "__main__":
    main()
