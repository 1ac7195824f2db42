"""ctypes binding of the C++ host facade (libkfusion_b200.so): kf::kinectfusion driven exactly
as an application drives the reference's class (kfusion/include/kinectfusion.h:31-73)."""
import ctypes as C
import os

import numpy as np

from .binding import Intrinsics, KfbError, KFB_MAX_LEVELS, load_library, Context

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_vp = C.c_void_p
BCAST_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.c_void_p)
COMPOSITE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class HostParams(C.Structure):
    _fields_ = [("pyramid_height", C.c_int), ("dfilter_dist", C.c_float), ("bfilter_kernel_size", C.c_int),
                ("bfilter_spatial_sigma", C.c_float), ("bfilter_color_sigma", C.c_float),
                ("icp_dist_threshold", C.c_float), ("icp_angle_threshold", C.c_float),
                ("icp_iter_count", C.c_int * KFB_MAX_LEVELS), ("volu_dims", C.c_int * 3),
                ("volu_range", C.c_float * 3), ("volu_pose", C.c_float * 12), ("volu_trun_dist", C.c_float),
                ("tsdf_max_weight", C.c_int), ("compat_icp_rows", C.c_int), ("compat_raycast_ts_sign", C.c_int),
                ("device", C.c_int), ("slab_z_begin", C.c_int), ("slab_z_end", C.c_int), ("shard_rank", C.c_int),
                ("shard_world", C.c_int)]


def host_library_path():
    return os.path.join(_HERE, "libkfusion_b200.so")


def load_host_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    load_library()  # dependency, loads libkfb200.so first (fails loudly if missing)
    path = host_library_path()
    if not os.path.exists(path):
        raise KfbError(f"{path} is missing: build it with __graft_entry__.build()")
    L = C.CDLL(path)
    sig = {
        "kfh_default_params": (None, [C.POINTER(HostParams), C.c_int]),
        "kfh_create": (_vp, [C.POINTER(Intrinsics), C.POINTER(HostParams)]),
        "kfh_last_error": (C.c_char_p, []),
        "kfh_destroy": (None, [_vp]),
        "kfh_reset": (None, [_vp]),
        "kfh_pipeline": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
        "kfh_frame_count": (C.c_int, [_vp]),
        "kfh_last_icp_us": (C.c_double, [_vp]),
        "kfh_num_poses": (C.c_int, [_vp]),
        "kfh_get_pose": (None, [_vp, C.c_int, _vp]),
        "kfh_context": (_vp, [_vp]),
        "kfh_render": (C.c_int, [_vp, C.c_int, _vp]),
        "kfh_extract_pointcloud": (C.c_long, [_vp, _vp, C.c_long]),
        "kfh_save_pointcloud": (C.c_int, [_vp, C.c_char_p]),
        "kfh_icp_solve": (C.c_int, [_vp, _vp]),
        "kfh_save_poses": (C.c_int, [_vp, C.c_char_p]),
        "kfh_save_volume": (C.c_int, [_vp, C.c_char_p]),
        "kfh_load_volume": (C.c_int, [_vp, C.c_char_p]),
        "kfh_read_intrinsics": (C.c_int, [C.c_char_p, _vp]),
        "kfh_icp_probe": (C.c_int, [_vp, _vp, _vp, C.c_int]),
        "kfh_sensor_open": (_vp, [C.c_char_p]),
        "kfh_sensor_close": (None, [_vp]),
        "kfh_sensor_info": (C.c_int, [_vp, _vp]),
        "kfh_sensor_get_frame": (C.c_int, [_vp, _vp, _vp]),
        "kfh_sensor_error": (C.c_char_p, [_vp]),
        "kfh_png_write_gray16": (C.c_int, [C.c_char_p, _vp, C.c_int, C.c_int]),
        "kfh_png_write_rgb8": (C.c_int, [C.c_char_p, _vp, C.c_int, C.c_int]),
        "kfh_set_shard_comm": (None, [_vp, BCAST_FN, COMPOSITE_FN, _vp]),
        "kfh_mailbox_open": (_vp, [C.c_char_p, C.c_int, C.c_int]),
        "kfh_mailbox_close": (None, [_vp]),
        "kfh_mailbox_exchange": (C.c_int, [_vp, _vp]),
        "kfh_set_pose_mailbox": (None, [_vp, _vp, COMPOSITE_FN, _vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def default_host_params(dims=512):
    p = HostParams()
    load_host_library().kfh_default_params(C.byref(p), int(dims))
    return p


def icp_probe(ctx, iters=(4, 5, 10)):
    """Wall-clock microseconds of every kfb_icp_step of one schedule, measured in C++ (diagnostic)."""
    it = (C.c_int * KFB_MAX_LEVELS)(*list(iters) + [0] * (KFB_MAX_LEVELS - len(iters)))
    out = np.zeros(64, np.float64)
    n = load_host_library().kfh_icp_probe(ctx.h, it, out.ctypes.data_as(_vp), 64)
    return out[:max(n, 0)]


def read_intrinsics(path):
    """Dataset intr.txt -> (fx, cx, fy, cy, depth scale), or None (depth_sensor.cpp:23-46)."""
    out = np.zeros(5, np.float32)
    rc = load_host_library().kfh_read_intrinsics(str(path).encode(), out.ctypes.data_as(_vp))
    return None if rc else out


class DatasetSensor:
    """The reference's dataset frame source (depth_sensor.cpp:11-46,186-196): color/*.png, depth/*.png (16-bit mm),
    intr.txt.  Iterating yields (bgr uint8 [h, w, 3], depth_mm float32 [h, w])."""

    def __init__(self, path):
        self.lib = load_host_library()
        self.h = self.lib.kfh_sensor_open(str(path).encode())
        if not self.h:
            raise FileNotFoundError(f"error: no camera! ({path})")
        info = np.zeros(8, np.float32)
        self.lib.kfh_sensor_info(self.h, info.ctypes.data_as(_vp))
        self.width, self.height = int(info[0]), int(info[1])
        self.fx, self.cx, self.fy, self.cy, self.scale = (float(v) for v in info[2:7])

    def frames_left(self):
        info = np.zeros(8, np.float32)
        self.lib.kfh_sensor_info(self.h, info.ctypes.data_as(_vp))
        return int(info[7])

    def get_frame(self):
        depth = np.empty((self.height, self.width), np.float32)
        bgr = np.empty((self.height, self.width, 3), np.uint8)
        rc = self.lib.kfh_sensor_get_frame(self.h, depth.ctypes.data_as(_vp), bgr.ctypes.data_as(_vp))
        if rc == 2:
            raise ValueError("frame size differs from the first colour image")
        return None if rc else (bgr, depth)

    def last_error(self):
        return self.lib.kfh_sensor_error(self.h).decode()

    def __iter__(self):
        while True:
            f = self.get_frame()
            if f is None:
                return
            yield f

    def close(self):
        if self.h:
            self.lib.kfh_sensor_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_png_gray16(path, img):
    a = np.ascontiguousarray(img, np.uint16)
    return load_host_library().kfh_png_write_gray16(str(path).encode(), a.ctypes.data_as(_vp), a.shape[1], a.shape[0]) == 0


def write_png_rgb8(path, img):
    a = np.ascontiguousarray(img, np.uint8)
    return load_host_library().kfh_png_write_rgb8(str(path).encode(), a.ctypes.data_as(_vp), a.shape[1], a.shape[0]) == 0


def icp_solve(sums27):
    s = np.ascontiguousarray(sums27, np.float64)
    x = np.zeros(6, np.float64)
    rc = load_host_library().kfh_icp_solve(s.ctypes.data_as(_vp), x.ctypes.data_as(_vp))
    return rc, x


class _BorrowedContext(Context):
    """A Context view over the kfb_ctx owned by a kf::kinectfusion (test hooks only)."""

    def __init__(self, handle, intr, params):  # noqa: D401 - no kfb_create here
        self.lib = load_library()
        self.h = _vp(handle)
        self.intr, self.params = intr, params

    def close(self):
        self.h = None


class NativePoseMailbox:
    """kf::PoseMailbox (kfusion/include/pose_mailbox.hpp): rank 0 must open before the others (barrier in between)."""

    def __init__(self, name, rank, world):
        self.lib = load_host_library()
        self.h = self.lib.kfh_mailbox_open(name.encode(), int(rank), int(world))
        if not self.h:
            raise KfbError("pose mailbox: " + self.lib.kfh_last_error().decode())

    def exchange(self, msg13):
        """msg13: contiguous float32[13] numpy array (in place)."""
        return self.lib.kfh_mailbox_exchange(self.h, msg13.ctypes.data_as(_vp))

    def close(self):
        if getattr(self, "h", None):
            self.lib.kfh_mailbox_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KinectFusion:
    """kf::kinectfusion(intr, params).pipeline(cmap, dmap) -- the public API a user calls."""

    def __init__(self, intr, params=None):
        self.lib = load_host_library()
        self.intr = intr
        self.params = params if params is not None else default_host_params(512)
        self.h = self.lib.kfh_create(C.byref(intr), C.byref(self.params))
        if not self.h:
            raise KfbError("kf::kinectfusion construction failed: " + self.lib.kfh_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.kfh_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pipeline(self, depth_mm):
        """depth_mm: float32 HxW millimetres (numpy, pageable or pinned)."""
        d = np.ascontiguousarray(depth_mm, np.float32)
        self._keep = d
        return self.lib.kfh_pipeline(self.h, d.ctypes.data_as(_vp), d.shape[1], d.shape[0])

    def pipeline_ptr(self, ptr, w, h):
        return self.lib.kfh_pipeline(self.h, ptr, w, h)

    def last_error(self):
        return self.lib.kfh_last_error().decode()

    def set_shard_comm(self, broadcast_pose, composite):
        """broadcast_pose(msg13: ctypes float*) -> int, composite() -> int (kf::ShardComm)."""
        self._cb = (BCAST_FN(lambda p, u: int(broadcast_pose(p))), COMPOSITE_FN(lambda u: int(composite())))
        self.lib.kfh_set_shard_comm(self.h, self._cb[0], self._cb[1], None)

    def set_pose_mailbox(self, mailbox, composite):
        """The native kf::PoseMailbox as broadcast_pose (no Python on the pose path); composite() -> int as above."""
        self._cb = (COMPOSITE_FN(lambda u: int(composite())),)
        self._mailbox = mailbox
        self.lib.kfh_set_pose_mailbox(self.h, mailbox.h, self._cb[0], None)

    def reset(self):
        self.lib.kfh_reset(self.h)

    def last_icp_us(self):
        return self.lib.kfh_last_icp_us(self.h)

    @property
    def frame_count(self):
        return self.lib.kfh_frame_count(self.h)

    def pose(self, idx=-1):
        p = np.empty(12, np.float32)
        self.lib.kfh_get_pose(self.h, idx, p.ctypes.data_as(_vp))
        return p

    def poses(self):
        return np.stack([self.pose(i) for i in range(self.lib.kfh_num_poses(self.h))])

    def context(self):
        from .binding import Params
        p = Params()
        load_library().kfb_default_params(C.byref(p), int(self.params.volu_dims[0]))
        for i in range(3):
            p.volu_dims[i] = self.params.volu_dims[i]
            p.volu_range[i] = self.params.volu_range[i]
        p.pyramid_height = self.params.pyramid_height
        p.slab_z_begin, p.slab_z_end = self.params.slab_z_begin, self.params.slab_z_end
        return _BorrowedContext(self.lib.kfh_context(self.h), self.intr, p)

    def render(self, normal=False):
        out = np.empty((self.intr.height, self.intr.width, 3), np.uint8)
        self.lib.kfh_render(self.h, int(normal), out.ctypes.data_as(_vp))
        return out

    def extract_pointcloud(self, cap=10_000_000):
        pts = np.empty((cap, 3), np.float32)
        n = self.lib.kfh_extract_pointcloud(self.h, pts.ctypes.data_as(_vp), cap)
        return pts[:min(n, cap)].copy()

    def save_pointcloud(self, path):
        self.lib.kfh_save_pointcloud(self.h, path.encode())

    def save_volume(self, path):
        return self.lib.kfh_save_volume(self.h, str(path).encode())

    def load_volume(self, path):
        return self.lib.kfh_load_volume(self.h, str(path).encode())

    def save_poses(self, path):
        return self.lib.kfh_save_poses(self.h, str(path).encode())
