"""Synthetic frame source (stands in for the reference's depth_sensor, kfusion/src/depth_sensor.cpp,
which is out of scope: device SDKs / PNG I/O).  Same contract: float32 depth in MILLIMETRES
(depth_sensor.cpp:192) plus intrinsics.  Analytic inside-out room + sphere + box and the looped
trajectory of SURVEY.md §8d, evaluated in float64 with numpy."""
import numpy as np


def trajectory_pose(k, period=300):
    """camera->world pose12 of frame k: t = (0.10 sin th, 0.05 sin 2th, 0.08 (1 - cos th)),
    yaw = 6 deg sin th, pitch = 3 deg sin 2th, th = 2 pi k / period."""
    th = 2.0 * np.pi * float(k) / float(period)
    yaw = (6.0 * np.pi / 180.0) * np.sin(th)
    pitch = (3.0 * np.pi / 180.0) * np.sin(2.0 * th)
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    R = np.array([[cy, sy * sp, sy * cp], [0.0, cp, -sp], [-sy, cy * sp, cy * cp]])
    t = np.array([0.10 * np.sin(th), 0.05 * np.sin(2.0 * th), 0.08 * (1.0 - np.cos(th))])
    return np.concatenate([R, t[:, None]], axis=1).astype(np.float32).reshape(12)


def render_depth_mm(pose12, width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5):
    """z-depth in mm, rounded to uint16 range, stored as float32."""
    P = np.asarray(pose12, np.float32).astype(np.float64).reshape(3, 4)
    R, o = P[:, :3], P[:, 3]
    u = (np.arange(width, dtype=np.float64) - np.float64(np.float32(cx))) / np.float64(np.float32(fx))
    v = (np.arange(height, dtype=np.float64) - np.float64(np.float32(cy))) / np.float64(np.float32(fy))
    U, V = np.meshgrid(u, v)
    d = [R[i, 0] * U + R[i, 1] * V + R[i, 2] * 1.0 for i in range(3)]
    best = np.full(U.shape, 1e30)

    def hit(t):
        nonlocal best
        ok = (t > 1e-9) & (t < best)
        best = np.where(ok, t, best)

    with np.errstate(divide="ignore", invalid="ignore"):
        hit(np.where(d[0] > 0, (1.3 - o[0]) / d[0], -1.0))
        hit(np.where(d[0] < 0, (-1.3 - o[0]) / d[0], -1.0))
        hit(np.where(d[1] > 0, (1.1 - o[1]) / d[1], -1.0))
        hit(np.where(d[1] < 0, (-1.1 - o[1]) / d[1], -1.0))
        hit(np.where(d[2] > 0, (3.1 - o[2]) / d[2], -1.0))
        # sphere
        sc, sr = np.array([0.35, 0.15, 1.9]), 0.35
        oc = o - sc
        a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
        b = 2.0 * (oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2])
        c = oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - sr * sr
        disc = b * b - 4.0 * a * c
        hit(np.where(disc >= 0, (-b - np.sqrt(np.maximum(disc, 0.0))) / (2.0 * a), -1.0))
        # box
        bmin, bmax = np.array([-0.8, 0.3, 1.6]), np.array([-0.2, 1.1, 2.2])
        t0 = np.full(U.shape, -1e30)
        t1 = np.full(U.shape, 1e30)
        ok = np.ones(U.shape, bool)
        for i in range(3):
            par = np.abs(d[i]) < 1e-12
            ok &= ~(par & ((o[i] < bmin[i]) | (o[i] > bmax[i])))
            ta, tb = (bmin[i] - o[i]) / d[i], (bmax[i] - o[i]) / d[i]
            lo, hi = np.minimum(ta, tb), np.maximum(ta, tb)
            t0 = np.where(par, t0, np.maximum(t0, lo))
            t1 = np.where(par, t1, np.minimum(t1, hi))
        hit(np.where(ok & (t0 <= t1), t0, -1.0))
    mm = np.where(best < 1e29, np.floor(best * 1000.0 + 0.5), 0.0)
    mm = np.where(mm > 65535.0, 0.0, mm)
    return mm.astype(np.float32)


def sequence(n_frames, intr, period=300, start=0):
    """List of (pose12, depth_mm) for frames start .. start+n_frames-1."""
    out = []
    for k in range(start, start + n_frames):
        p = trajectory_pose(k, period)
        out.append((p, render_depth_mm(p, intr.width, intr.height, intr.fx, intr.fy, intr.cx, intr.cy)))
    return out
