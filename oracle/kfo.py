"""ctypes binding of the CPU oracle (oracle/libkf_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkf_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, seconds).  Safe to call repeatedly."""
    src = os.path.join(_HERE, "kf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libkf_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Intr(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float)]

    def level(self, l):
        out = Intr()
        lib().kfo_level_intrinsics(C.byref(self), l, C.byref(out))
        return out


class VolumeDesc(C.Structure):
    _fields_ = [("dims", C.c_int * 3), ("range", C.c_float * 3), ("voxel_size", C.c_float * 3),
                ("trunc_dist", C.c_float), ("max_weight", C.c_int)]


class Params(C.Structure):
    _fields_ = [("pyramid_height", C.c_int), ("dfilter_dist", C.c_float), ("bfilter_kernel_size", C.c_int),
                ("bfilter_spatial_sigma", C.c_float), ("bfilter_color_sigma", C.c_float),
                ("icp_dist_threshold", C.c_float), ("icp_angle_threshold", C.c_float),
                ("icp_iter_count", C.c_int * 8), ("volu_range", C.c_float * 3), ("volu_pose", C.c_float * 12),
                ("volu_trun_dist", C.c_float), ("volu_dims", C.c_int * 3), ("tsdf_max_weight", C.c_int),
                ("compat_icp_rows", C.c_int), ("compat_raycast_ts_sign", C.c_int)]


_lib = None
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_sp = np.ctypeslib.ndpointer(dtype=np.int16, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i64p = C.POINTER(C.c_int64)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    IP, VP, PP = C.POINTER(Intr), C.POINTER(VolumeDesc), C.POINTER(Params)
    sig = {
        "kfo_level_intrinsics": (None, [IP, C.c_int, IP]),
        "kfo_pyrdown": (None, [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_int]),
        "kfo_bilateral": (None, [_fp, C.c_int, C.c_int, _fp, C.c_int, C.c_float, C.c_float]),
        "kfo_truncate": (None, [_fp, C.c_int, C.c_int, C.c_float]),
        "kfo_vertex_map": (None, [_fp, IP, _fp]),
        "kfo_normal_map": (None, [_fp, C.c_int, C.c_int, _fp]),
        "kfo_resize_maps": (None, [_fp, _fp, C.c_int, C.c_int, _fp, _fp]),
        "kfo_icp_accumulate": (None, [_fp, _fp, _fp, _fp, IP, _fp, C.c_float, C.c_float, C.c_int, _dp, _i64p]),
        "kfo_icp_solve": (C.c_int, [_dp, _dp]),
        "kfo_pose_apply_increment": (None, [_fp, _dp]),
        "kfo_pose_identity": (None, [_fp]),
        "kfo_pose_mul": (None, [_fp, _fp, _fp]),
        "kfo_pose_inv": (None, [_fp, _fp]),
        "kfo_rot_inv": (None, [_fp, _fp]),
        "kfo_rodrigues": (None, [_fp, _fp]),
        "kfo_integrate": (None, [_sp, VP, _fp, _fp, IP, C.c_int, C.c_int, _i64p]),
        "kfo_raycast": (None, [_sp, VP, _fp, _fp, IP, _fp, _fp, _i64p, C.c_int]),
        "kfo_raycast_slab": (None, [_sp, VP, _fp, _fp, IP, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "kfo_extract_points": (C.c_int64, [_sp, VP, _fp, _fp, C.c_int64]),
        "kfo_render_phong": (None, [_fp, _fp, C.c_int, C.c_int, _fp, _bp]),
        "kfo_render_normals": (None, [_fp, C.c_int, C.c_int, _bp]),
        "kfo_trajectory_pose": (None, [C.c_int, C.c_int, _fp]),
        "kfo_render_depth_mm": (None, [_fp, IP, _fp]),
        "kfo_fill_const_depth_mm": (None, [C.c_int, C.c_int, C.c_float, _fp]),
        "kfo_default_params": (None, [PP, C.c_int]),
        "kfo_kinfu_create": (C.c_void_p, [IP, PP]),
        "kfo_kinfu_destroy": (None, [C.c_void_p]),
        "kfo_kinfu_reset": (None, [C.c_void_p]),
        "kfo_kinfu_pipeline": (C.c_int, [C.c_void_p, _fp]),
        "kfo_kinfu_frame_count": (C.c_int, [C.c_void_p]),
        "kfo_kinfu_num_poses": (C.c_int, [C.c_void_p]),
        "kfo_kinfu_get_pose": (None, [C.c_void_p, C.c_int, _fp]),
        "kfo_kinfu_volume": (C.POINTER(C.c_int16), [C.c_void_p]),
        "kfo_kinfu_cur_depth": (C.POINTER(C.c_float), [C.c_void_p, C.c_int]),
        "kfo_kinfu_cur_vmap": (C.POINTER(C.c_float), [C.c_void_p, C.c_int]),
        "kfo_kinfu_cur_nmap": (C.POINTER(C.c_float), [C.c_void_p, C.c_int]),
        "kfo_kinfu_prev_vmap": (C.POINTER(C.c_float), [C.c_void_p, C.c_int]),
        "kfo_kinfu_prev_nmap": (C.POINTER(C.c_float), [C.c_void_p, C.c_int]),
        "kfo_kinfu_last_updated": (C.c_int64, [C.c_void_p]),
        "kfo_kinfu_last_raysteps": (C.c_int64, [C.c_void_p]),
        "kfo_kinfu_last_times": (None, [C.c_void_p, _dp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


# ---------------------------------------------------------------- helpers
def intr(width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5):
    return Intr(width, height, fx, fy, cx, cy)


SENSORS = {
    "kinect1": dict(width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5),
    "kinect2": dict(width=512, height=424, fx=365.5, fy=365.5, cx=255.5, cy=211.5),
    "realsense720": dict(width=1280, height=720, fx=920.0, fy=920.0, cx=639.5, cy=359.5),
}


def volume_desc(dims=512, rng=3.0, trunc=None, max_weight=64):
    vd = VolumeDesc()
    d = (dims,) * 3 if np.isscalar(dims) else tuple(dims)
    for i in range(3):
        vd.dims[i] = int(d[i])
        vd.range[i] = np.float32(rng)
        vd.voxel_size[i] = np.float32(rng) / np.float32(d[i])
    vd.trunc_dist = np.float32(2.1) * np.float32(rng) / np.float32(d[0]) if trunc is None else np.float32(trunc)
    vd.max_weight = max_weight
    return vd


def default_params(dims=512):
    p = Params()
    lib().kfo_default_params(C.byref(p), int(dims))
    return p


def identity():
    p = np.zeros(12, np.float32)
    p[0] = p[5] = p[10] = 1
    return p


def pose_mul(a, b):
    o = np.empty(12, np.float32)
    lib().kfo_pose_mul(np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32), o)
    return o


def pose_inv(a):
    o = np.empty(12, np.float32)
    lib().kfo_pose_inv(np.ascontiguousarray(a, np.float32), o)
    return o


def rot_inv(a):
    o = np.empty(9, np.float32)
    lib().kfo_rot_inv(np.ascontiguousarray(a, np.float32), o)
    return o


def trajectory_pose(k, period=300):
    p = np.empty(12, np.float32)
    lib().kfo_trajectory_pose(int(k), int(period), p)
    return p


def render_depth_mm(pose12, K):
    d = np.empty((K.height, K.width), np.float32)
    lib().kfo_render_depth_mm(np.ascontiguousarray(pose12, np.float32), C.byref(K), d)
    return d


def pyrdown(src):
    h, w = src.shape
    dst = np.empty((h >> 1, w >> 1), np.float32)
    lib().kfo_pyrdown(np.ascontiguousarray(src, np.float32), w, h, dst, w >> 1, h >> 1)
    return dst


def bilateral(src, ksize=5, sigma_color=10.0, sigma_space=10.0):
    h, w = src.shape
    dst = np.empty_like(src, dtype=np.float32)
    lib().kfo_bilateral(np.ascontiguousarray(src, np.float32), w, h, dst, ksize, sigma_color, sigma_space)
    return dst


def truncate(d, max_dist=5.0):
    out = np.array(d, dtype=np.float32, copy=True)
    h, w = out.shape
    lib().kfo_truncate(out, w, h, max_dist)
    return out


def vertex_map(d, K):
    v = np.empty((K.height, K.width, 3), np.float32)
    lib().kfo_vertex_map(np.ascontiguousarray(d, np.float32), C.byref(K), v)
    return v


def normal_map(v):
    h, w, _ = v.shape
    n = np.empty_like(v)
    lib().kfo_normal_map(np.ascontiguousarray(v, np.float32), w, h, n)
    return n


def resize_maps(vbig, nbig):
    h, w, _ = vbig.shape
    vs = np.empty((h >> 1, w >> 1, 3), np.float32)
    ns = np.empty((h >> 1, w >> 1, 3), np.float32)
    lib().kfo_resize_maps(np.ascontiguousarray(vbig), np.ascontiguousarray(nbig), w, h, vs, ns)
    return vs, ns


def frontend(depth_mm, K, levels=3, ksize=5, sc=10.0, ss=10.0, max_dist=5.0):
    """pyrDown chain + bilateral + truncate + vertex + normal, like kinectfusion.cpp:48-76."""
    raw = [np.ascontiguousarray(depth_mm, np.float32)]
    for l in range(1, levels):
        raw.append(pyrdown(raw[-1]))
    out = []
    for l in range(levels):
        d = truncate(bilateral(raw[l], ksize, sc, ss), max_dist)
        Kl = K.level(l)
        v = vertex_map(d, Kl)
        out.append((d, v, normal_map(v)))
    return out


def icp_accumulate(cur_v, cur_n, pre_v, pre_n, K, pose12, dist=0.015, sine=0.5, compat_rows=1):
    out = np.empty(27, np.float64)
    cnt = C.c_int64(0)
    lib().kfo_icp_accumulate(np.ascontiguousarray(cur_v, np.float32), np.ascontiguousarray(cur_n, np.float32),
                             np.ascontiguousarray(pre_v, np.float32), np.ascontiguousarray(pre_n, np.float32),
                             C.byref(K), np.ascontiguousarray(pose12, np.float32), dist, sine, compat_rows, out,
                             C.byref(cnt))
    return out, cnt.value


def icp_solve(ab27):
    x = np.zeros(6, np.float64)
    rc = lib().kfo_icp_solve(np.ascontiguousarray(ab27, np.float64), x)
    return rc, x


def pose_apply_increment(pose12, x6):
    p = np.array(pose12, dtype=np.float32, copy=True)
    lib().kfo_pose_apply_increment(p, np.ascontiguousarray(x6, np.float64))
    return p


def new_volume(vd):
    return np.zeros((vd.dims[2], vd.dims[1], vd.dims[0], 2), np.int16)


def integrate(vol, vd, vol2cam, depth_m, K, z_begin=1, z_end=None):
    upd = C.c_int64(0)
    lib().kfo_integrate(vol.reshape(-1), C.byref(vd), np.ascontiguousarray(vol2cam, np.float32),
                        np.ascontiguousarray(depth_m, np.float32), C.byref(K), z_begin,
                        vd.dims[2] if z_end is None else z_end, C.byref(upd))
    return upd.value


def raycast(vol, vd, cam2vol, K, compat_ts_sign=1):
    v = np.empty((K.height, K.width, 3), np.float32)
    n = np.empty((K.height, K.width, 3), np.float32)
    steps = C.c_int64(0)
    lib().kfo_raycast(vol.reshape(-1), C.byref(vd), np.ascontiguousarray(cam2vol, np.float32), rot_inv(cam2vol),
                      C.byref(K), v, n, C.byref(steps), compat_ts_sign)
    return v, n, steps.value


def raycast_slab(vol_slab, vd, cam2vol, K, zs0, zs1, zo0, zo1, compat_ts_sign=1):
    """vol_slab holds planes [zs0, zs1) of the volume; returns (vmap, nmap, key)."""
    v = np.empty((K.height, K.width, 3), np.float32)
    n = np.empty((K.height, K.width, 3), np.float32)
    key = np.empty((K.height, K.width), np.float32)
    lib().kfo_raycast_slab(np.ascontiguousarray(vol_slab).reshape(-1), C.byref(vd), np.ascontiguousarray(cam2vol, np.float32),
                           rot_inv(cam2vol), C.byref(K), v, n, key, zs0, zs1, zo0, zo1, compat_ts_sign)
    return v, n, key


def extract_points(vol, vd, volpose, cap=10_000_000):
    pts = np.empty((cap, 3), np.float32)
    n = lib().kfo_extract_points(vol.reshape(-1), C.byref(vd), np.ascontiguousarray(volpose, np.float32), pts, cap)
    return pts[:n].copy()


def render_phong(v, n, eye):
    h, w, _ = v.shape
    out = np.zeros((h, w, 3), np.uint8)
    lib().kfo_render_phong(np.ascontiguousarray(v), np.ascontiguousarray(n), w, h,
                           np.ascontiguousarray(eye, np.float32), out)
    return out


def render_normals(n):
    h, w, _ = n.shape
    out = np.zeros((h, w, 3), np.uint8)
    lib().kfo_render_normals(np.ascontiguousarray(n), w, h, out)
    return out


class Kinfu:
    """The whole reference pipeline (kf::kinectfusion) on the CPU."""

    def __init__(self, K, params):
        self.K, self.p = K, params
        self.h = lib().kfo_kinfu_create(C.byref(K), C.byref(params))

    def __del__(self):
        if getattr(self, "h", None):
            lib().kfo_kinfu_destroy(self.h)
            self.h = None

    def pipeline(self, depth_mm):
        return lib().kfo_kinfu_pipeline(self.h, np.ascontiguousarray(depth_mm, np.float32))

    def reset(self):
        lib().kfo_kinfu_reset(self.h)

    @property
    def frame_count(self):
        return lib().kfo_kinfu_frame_count(self.h)

    def poses(self):
        n = lib().kfo_kinfu_num_poses(self.h)
        out = np.empty((n, 12), np.float32)
        for i in range(n):
            lib().kfo_kinfu_get_pose(self.h, i, out[i])
        return out

    def pose(self):
        p = np.empty(12, np.float32)
        lib().kfo_kinfu_get_pose(self.h, -1, p)
        return p

    def volume(self):
        d = self.p.volu_dims
        ptr = lib().kfo_kinfu_volume(self.h)
        return np.ctypeslib.as_array(ptr, shape=(d[2], d[1], d[0], 2))

    def _map(self, fn, level, ch):
        Kl = self.K.level(level)
        ptr = fn(self.h, level)
        shape = (Kl.height, Kl.width, 3) if ch == 3 else (Kl.height, Kl.width)
        return np.ctypeslib.as_array(ptr, shape=shape)

    def cur_depth(self, l=0):
        return self._map(lib().kfo_kinfu_cur_depth, l, 1)

    def cur_vmap(self, l=0):
        return self._map(lib().kfo_kinfu_cur_vmap, l, 3)

    def cur_nmap(self, l=0):
        return self._map(lib().kfo_kinfu_cur_nmap, l, 3)

    def prev_vmap(self, l=0):
        return self._map(lib().kfo_kinfu_prev_vmap, l, 3)

    def prev_nmap(self, l=0):
        return self._map(lib().kfo_kinfu_prev_nmap, l, 3)

    @property
    def last_updated(self):
        return lib().kfo_kinfu_last_updated(self.h)

    @property
    def last_raysteps(self):
        return lib().kfo_kinfu_last_raysteps(self.h)

    def last_times(self):
        t = np.zeros(4, np.float64)
        lib().kfo_kinfu_last_times(self.h, t)
        return t
