"""ctypes binding of include/kfb200.h (libkfb200.so).  Harness only -- no compute here."""
import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KFB_MAX_LEVELS = 8
FRAME_CUR, FRAME_PREV = 0, 1


class KfbError(RuntimeError):
    pass


class Intrinsics(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float)]

    def level(self, l):
        out = Intrinsics()
        load_library().kfb_level_intrinsics(C.byref(self), int(l), C.byref(out))
        return out


class Params(C.Structure):
    _fields_ = [("pyramid_height", C.c_int), ("dfilter_dist", C.c_float), ("bfilter_kernel_size", C.c_int),
                ("bfilter_spatial_sigma", C.c_float), ("bfilter_color_sigma", C.c_float),
                ("icp_dist_threshold", C.c_float), ("icp_angle_threshold", C.c_float),
                ("icp_iter_count", C.c_int * KFB_MAX_LEVELS), ("volu_dims", C.c_int * 3),
                ("volu_range", C.c_float * 3), ("volu_trun_dist", C.c_float), ("tsdf_max_weight", C.c_int),
                ("compat_icp_rows", C.c_int), ("compat_raycast_ts_sign", C.c_int),
                ("slab_z_begin", C.c_int), ("slab_z_end", C.c_int)]


SENSORS = {
    "kinect1": dict(width=640, height=480, fx=525.0, fy=525.0, cx=319.5, cy=239.5),
    "kinect2": dict(width=512, height=424, fx=365.5, fy=365.5, cx=255.5, cy=211.5),
    "realsense720": dict(width=1280, height=720, fx=920.0, fy=920.0, cx=639.5, cy=359.5),
}


def library_path():
    return os.path.join(_HERE, "libkfb200.so")


def exported_symbols(header=None):
    """Every function name include/kfb200.h declares (used by the symbol-export test)."""
    header = header or os.path.join(os.path.dirname(_HERE), "include", "kfb200.h")
    txt = re.sub(r"/\*.*?\*/", "", open(header).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(kfb_[a-z0-9_]+)\s*\(", txt)))


_fp = C.POINTER(C.c_float)
_vp = C.c_void_p


def load_library():
    """Load libkfb200.so.  Fails loudly if the CUDA extension has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise KfbError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(path)
    IP, PP = C.POINTER(Intrinsics), C.POINTER(Params)
    sig = {
        "kfb_default_params": (None, [PP, C.c_int]),
        "kfb_create": (C.c_int, [IP, PP, C.c_int, C.POINTER(_vp)]),
        "kfb_destroy": (None, [_vp]),
        "kfb_last_error_string": (C.c_char_p, [_vp]),
        "kfb_synchronize": (C.c_int, [_vp]),
        "kfb_device_count": (C.c_int, []),
        "kfb_set_stream": (C.c_int, [_vp, _vp]),
        "kfb_composite_mask": (C.c_int, [_vp, _vp]),
        "kfb_ipc_export": (C.c_int, [_vp, C.c_int, _vp]),
        "kfb_shard_attach": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
        "kfb_shard_composite": (C.c_int, [_vp]),
        "kfb_shard_attached": (C.c_int, [_vp]),
        "kfb_shard_detach": (C.c_int, [_vp]),
        "kfb_reset_volume": (C.c_int, [_vp]),
        "kfb_reset_frames": (C.c_int, [_vp]),
        "kfb_upload_depth_mm": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
        "kfb_upload_depth_mm_u16": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
        "kfb_upload_wait": (C.c_int, [_vp]),
        "kfb_frontend": (C.c_int, [_vp]),
        "kfb_swap_frames": (C.c_int, [_vp]),
        "kfb_icp_accumulate": (C.c_int, [_vp, C.c_int, _vp, _vp]),
        "kfb_icp_begin": (C.c_int, [_vp, _vp]),
        "kfb_icp_step": (C.c_int, [_vp, _vp, _vp]),
        "kfb_icp_end": (C.c_int, [_vp]),
        "kfb_integrate": (C.c_int, [_vp, _vp, C.POINTER(C.c_uint64)]),
        "kfb_raycast": (C.c_int, [_vp, _vp, _vp]),
        "kfb_integrate_plane_histogram": (C.c_int, [_vp, _vp, _vp]),
        "kfb_model_pyramid": (C.c_int, [_vp]),
        "kfb_extract_points": (C.c_int, [_vp, _vp, _vp, C.c_size_t, C.POINTER(C.c_size_t)]),
        "kfb_render_phong": (C.c_int, [_vp, _vp, _vp]),
        "kfb_render_normals": (C.c_int, [_vp, _vp]),
        "kfb_download_depth": (C.c_int, [_vp, C.c_int, _vp]),
        "kfb_upload_depth_m": (C.c_int, [_vp, C.c_int, _vp]),
        "kfb_download_raw_depth": (C.c_int, [_vp, C.c_int, _vp]),
        "kfb_download_maps": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
        "kfb_upload_maps": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp]),
        "kfb_download_volume": (C.c_int, [_vp, _vp]),
        "kfb_upload_volume": (C.c_int, [_vp, _vp]),
        "kfb_volume_voxels": (C.c_size_t, [_vp]),
        "kfb_level_intrinsics": (None, [IP, C.c_int, IP]),
        "kfb_event_record": (C.c_int, [_vp, C.c_int]),
        "kfb_event_elapsed_ms": (C.c_int, [_vp, C.c_int, C.c_int, _fp]),
        "kfb_set_profiling": (C.c_int, [_vp, C.c_int]),
        "kfb_launch_count": (C.c_uint64, [_vp]),
        "kfb_device_ptr": (_vp, [_vp, C.c_int]),
        "kfb_stream": (_vp, [_vp]),
        "kfb_debug_icp_ring": (None, [_vp, _vp]),
        "kfb_icp_fallback_count": (C.c_uint64, [_vp]),
        "kfb_icp_mispredict_count": (C.c_uint64, [_vp]),
        "kfb_debug_integrate_counts": (None, [_vp, _vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _LIB = L
    return L


def default_params(dims=512):
    p = Params()
    load_library().kfb_default_params(C.byref(p), int(dims))
    return p


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    return a.ctypes.data_as(_vp)


class Context:
    """Thin RAII wrapper of kfb_ctx.  Every method maps 1:1 onto a C-ABI entry point."""

    def __init__(self, intr, params, device=0):
        self.lib = load_library()
        self.intr, self.params = intr, params
        h = _vp()
        rc = self.lib.kfb_create(C.byref(intr), C.byref(params), int(device), C.byref(h))
        self.h = h
        if rc != 0:
            msg = self.lib.kfb_last_error_string(h).decode() if h else "invalid arguments"
            if h:
                self.lib.kfb_destroy(h)
            self.h = None
            raise KfbError(f"kfb_create failed (rc={rc}): {msg}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.kfb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise KfbError(f"kfb call failed (rc={rc}): {self.lib.kfb_last_error_string(self.h).decode()}")

    # ---- pipeline stages
    def reset_volume(self):
        self._ck(self.lib.kfb_reset_volume(self.h))

    def reset_frames(self):
        self._ck(self.lib.kfb_reset_frames(self.h))

    def upload_depth_mm(self, depth_mm):
        d = _f32(depth_mm)
        self._keep = d
        self._ck(self.lib.kfb_upload_depth_mm(self.h, _ptr(d), d.shape[1], d.shape[0]))

    def upload_depth_mm_u16(self, depth_mm_u16):
        d = np.ascontiguousarray(depth_mm_u16, dtype=np.uint16)
        self._keep = d
        self._ck(self.lib.kfb_upload_depth_mm_u16(self.h, _ptr(d), d.shape[1], d.shape[0]))

    def upload_depth_mm_ptr(self, ptr, w, h):
        self._ck(self.lib.kfb_upload_depth_mm(self.h, ptr, w, h))

    def frontend(self):
        self._ck(self.lib.kfb_frontend(self.h))

    def swap_frames(self):
        self._ck(self.lib.kfb_swap_frames(self.h))

    def icp_accumulate(self, level, pose12):
        p = _f32(pose12)
        out = np.empty(27, np.float64)
        self._ck(self.lib.kfb_icp_accumulate(self.h, int(level), _ptr(p), _ptr(out)))
        return out

    def icp_begin(self, iters_per_level):
        it = (C.c_int * KFB_MAX_LEVELS)(*list(iters_per_level) + [0] * (KFB_MAX_LEVELS - len(iters_per_level)))
        self._ck(self.lib.kfb_icp_begin(self.h, it))

    def icp_step(self, pose12):
        p = _f32(pose12)
        out = np.empty(27, np.float64)
        self._ck(self.lib.kfb_icp_step(self.h, _ptr(p), _ptr(out)))
        return out

    def icp_end(self):
        self._ck(self.lib.kfb_icp_end(self.h))

    def integrate(self, vol2cam12, count=False):
        p = _f32(vol2cam12)
        if count:
            n = C.c_uint64(0)
            self._ck(self.lib.kfb_integrate(self.h, _ptr(p), C.byref(n)))
            return n.value
        self._ck(self.lib.kfb_integrate(self.h, _ptr(p), None))
        return None

    def plane_histogram(self, vol2cam12):
        p = _f32(vol2cam12)
        out = np.zeros(int(self.params.volu_dims[2]), np.uint32)
        self._ck(self.lib.kfb_integrate_plane_histogram(self.h, _ptr(p), _ptr(out)))
        return out

    def raycast(self, cam2vol12, rinv9):
        p, r = _f32(cam2vol12), _f32(rinv9)
        self._ck(self.lib.kfb_raycast(self.h, _ptr(p), _ptr(r)))

    def set_stream(self, cuda_stream):
        self._ck(self.lib.kfb_set_stream(self.h, _vp(int(cuda_stream))))

    def composite_mask(self, min_key_ptr):
        self._ck(self.lib.kfb_composite_mask(self.h, _vp(int(min_key_ptr))))

    def ipc_export(self):
        """256 bytes: the four CUDA IPC handles (keys, maps slot 0, maps slot 1, flag) of this context."""
        buf = np.zeros(4 * 64, np.uint8)
        for w in range(4):
            self._ck(self.lib.kfb_ipc_export(self.h, w, _vp(buf.ctypes.data + 64 * w)))
        return buf

    def shard_attach(self, rank, world, handles):
        h = np.ascontiguousarray(handles, np.uint8)
        assert h.size == world * 256
        self._ck(self.lib.kfb_shard_attach(self.h, int(rank), int(world), _ptr(h)))

    def shard_detach(self):
        self._ck(self.lib.kfb_shard_detach(self.h))

    def shard_composite(self):
        self._ck(self.lib.kfb_shard_composite(self.h))

    def model_pyramid(self):
        self._ck(self.lib.kfb_model_pyramid(self.h))

    def extract_points(self, volpose12, cap=10_000_000):
        p = _f32(volpose12)
        out = np.empty((cap, 3), np.float32)
        n = C.c_size_t(0)
        self._ck(self.lib.kfb_extract_points(self.h, _ptr(p), _ptr(out), cap, C.byref(n)))
        return out[:n.value].copy()

    def render_phong(self, eye3):
        e = _f32(eye3)
        out = np.empty((self.intr.height, self.intr.width, 3), np.uint8)
        self._ck(self.lib.kfb_render_phong(self.h, _ptr(e), _ptr(out)))
        return out

    def render_normals(self):
        out = np.empty((self.intr.height, self.intr.width, 3), np.uint8)
        self._ck(self.lib.kfb_render_normals(self.h, _ptr(out)))
        return out

    # ---- hooks
    def _shape(self, level):
        k = self.intr.level(level)
        return k.height, k.width

    def download_depth(self, level=0):
        out = np.empty(self._shape(level), np.float32)
        self._ck(self.lib.kfb_download_depth(self.h, level, _ptr(out)))
        return out

    def download_raw_depth(self, level=0):
        out = np.empty(self._shape(level), np.float32)
        self._ck(self.lib.kfb_download_raw_depth(self.h, level, _ptr(out)))
        return out

    def upload_depth_m(self, level, depth_m):
        d = _f32(depth_m)
        assert d.shape == self._shape(level)
        self._ck(self.lib.kfb_upload_depth_m(self.h, level, _ptr(d)))

    def download_maps(self, frame, level=0):
        h, w = self._shape(level)
        v = np.empty((h, w, 3), np.float32)
        n = np.empty((h, w, 3), np.float32)
        self._ck(self.lib.kfb_download_maps(self.h, frame, level, _ptr(v), _ptr(n)))
        return v, n

    def upload_maps(self, frame, level, v, n):
        v, n = _f32(v), _f32(n)
        self._ck(self.lib.kfb_upload_maps(self.h, frame, level, _ptr(v), _ptr(n)))

    def volume_voxels(self):
        return self.lib.kfb_volume_voxels(self.h)

    def download_volume(self):
        d = self.params.volu_dims
        planes = self.volume_voxels() // (d[0] * d[1])
        out = np.empty((planes, d[1], d[0], 2), np.int16)
        self._ck(self.lib.kfb_download_volume(self.h, _ptr(out)))
        return out

    def upload_volume(self, vol):
        v = np.ascontiguousarray(vol, dtype=np.int16)
        assert v.size == 2 * self.volume_voxels()
        self._ck(self.lib.kfb_upload_volume(self.h, _ptr(v)))

    # ---- measurement
    def synchronize(self):
        self._ck(self.lib.kfb_synchronize(self.h))

    def event_record(self, slot):
        self._ck(self.lib.kfb_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float(0)
        self._ck(self.lib.kfb_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def set_profiling(self, on):
        self._ck(self.lib.kfb_set_profiling(self.h, int(bool(on))))

    def launch_count(self):
        return self.lib.kfb_launch_count(self.h)

    def device_ptr(self, which):
        return self.lib.kfb_device_ptr(self.h, which)

    def debug_icp_ring(self):
        out = np.zeros((32, 4), np.uint64)
        self.lib.kfb_debug_icp_ring(self.h, _ptr(out))
        return out

    def integrate_counts(self):
        """{updated, quads_loaded, quads_stored, stream_items, general_items} of the last integrate(count=True)."""
        out = np.zeros(6, np.uint64)
        self.lib.kfb_debug_integrate_counts(self.h, _ptr(out))
        return dict(updated=int(out[0]), quads_loaded=int(out[1]), quads_stored=int(out[2]), stream_items=int(out[3]), general_items=int(out[4]))

    def icp_fallback_count(self):
        return int(self.lib.kfb_icp_fallback_count(self.h))

    def icp_mispredict_count(self):
        return int(self.lib.kfb_icp_mispredict_count(self.h))

    def stream(self):
        return self.lib.kfb_stream(self.h)
