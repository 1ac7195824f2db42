"""ctypes binding of oracle/_ref/libkf_ref.so: the reference's own CUDA kernels, compiled
unmodified for sm_100a (oracle/Makefile).  TEST INFRASTRUCTURE ONLY -- the A/B checker on a
B200 and the `--impl reference` baseline of bench.py.  Needs a GPU at call time."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libkf_ref.so")
_lib = None
_vp = C.c_void_p


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_volume_create.restype = _vp
        L.ref_volume_create.argtypes = [_vp, _vp, C.c_float]
        L.ref_volume_destroy.argtypes = [_vp]
        L.ref_volume_upload.argtypes = [_vp, _vp]
        L.ref_volume_download.argtypes = [_vp, _vp]
        L.ref_integrate.restype = C.c_float
        L.ref_integrate.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int]
        L.ref_raycast.restype = C.c_float
        L.ref_raycast.argtypes = [_vp, _vp, _vp, C.c_int, C.c_int] + [C.c_float] * 4 + [_vp, _vp, C.c_int]
        L.ref_extract_points.restype = C.c_long
        L.ref_extract_points.argtypes = [_vp, _vp, _vp, C.c_long]
        L.ref_depth_truncation.argtypes = [_vp, C.c_int, C.c_int, C.c_float]
        L.ref_vertex_normal.argtypes = [_vp, C.c_int, C.c_int] + [C.c_float] * 4 + [_vp, _vp]
        L.ref_resize_maps.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, _vp]
        L.ref_render.argtypes = [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp]
        L.ref_rigid_icp.restype = C.c_float
        L.ref_rigid_icp.argtypes = [_vp] * 4 + [C.c_int, C.c_int] + [C.c_float] * 4 + [_vp, C.c_float, C.c_float, _vp, _vp, C.c_int]
        L.ref_last_cuda_error.restype = C.c_int
        L.ref_kinfu_create.restype = _vp
        L.ref_kinfu_create.argtypes = [C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int, C.c_float, _vp]
        L.ref_kinfu_destroy.argtypes = [_vp]
        L.ref_kinfu_reset.argtypes = [_vp]
        L.ref_kinfu_get_pose.argtypes = [_vp, _vp]
        L.ref_kinfu_pipeline.restype = C.c_int
        L.ref_kinfu_pipeline.argtypes = [_vp, _vp]
        L.ref_kinfu_sync_from.argtypes = [_vp, _vp, _vp, _vp, _vp, C.c_int]
        L.ref_event_ms.restype = C.c_float
        L.ref_event_ms.argtypes = [C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_vp)


def _f(a):
    return np.ascontiguousarray(a, np.float32)


class RefVolume:
    def __init__(self, dims, rng=3.0, trunc=None):
        d = (dims,) * 3 if np.isscalar(dims) else tuple(dims)
        self.dims = np.array(d, np.int32)
        self.range = np.array([rng] * 3, np.float32)
        self.trunc = np.float32(2.1) * np.float32(rng) / np.float32(d[0]) if trunc is None else np.float32(trunc)
        self.h = lib().ref_volume_create(_p(self.dims), _p(self.range), float(self.trunc))
        if not self.h:
            raise RuntimeError("reference volume allocation failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_volume_destroy(self.h)
            self.h = None

    def upload(self, vol_pairs):
        v = np.ascontiguousarray(vol_pairs, np.int16)
        lib().ref_volume_upload(self.h, _p(v))

    def download(self):
        out = np.empty((self.dims[2], self.dims[1], self.dims[0], 2), np.int16)
        lib().ref_volume_download(self.h, _p(out))
        return out

    def integrate(self, vol2cam12, depth_m, K, reps=1):
        d, p = _f(depth_m), _f(vol2cam12)
        return lib().ref_integrate(self.h, _p(p), _p(d), K.width, K.height, K.fx, K.fy, K.cx, K.cy, reps)

    def raycast(self, cam2vol12, rinv9, K, reps=1):
        p, r = _f(cam2vol12), _f(rinv9)
        v = np.empty((K.height, K.width, 3), np.float32)
        n = np.empty((K.height, K.width, 3), np.float32)
        ms = lib().ref_raycast(self.h, _p(p), _p(r), K.width, K.height, K.fx, K.fy, K.cx, K.cy, _p(v), _p(n), reps)
        return v, n, ms

    def extract_points(self, volpose12, cap=10_000_000):
        p = _f(volpose12)
        pts = np.empty((cap, 3), np.float32)
        n = lib().ref_extract_points(self.h, _p(p), _p(pts), cap)
        return pts[:n].copy()


class RefKinfu:
    """The reference's whole frame loop over its own kernels (ref_harness.cu: ref_kinfu_*)."""

    def __init__(self, K, dims, volpose12, rng=3.0):
        vp = _f(volpose12)
        self.h = lib().ref_kinfu_create(K.width, K.height, K.fx, K.fy, K.cx, K.cy, int(dims), float(rng), _p(vp))
        if not self.h:
            raise RuntimeError("reference pipeline allocation failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_kinfu_destroy(self.h)
            self.h = None

    def reset(self):
        lib().ref_kinfu_reset(self.h)

    def pipeline_ptr(self, ptr):
        return lib().ref_kinfu_pipeline(self.h, _vp(int(ptr)))

    def pipeline(self, depth_mm):
        d = _f(depth_mm)
        return lib().ref_kinfu_pipeline(self.h, _p(d))

    def pose(self):
        p = np.empty(12, np.float32)
        lib().ref_kinfu_get_pose(self.h, _p(p))
        return p

    def sync_from(self, vol_packed_dev, vmap4_dev, nmap4_dev, pose12, frame_count):
        """Adopt another pipeline's state (device pointers, same process): packed 4-byte voxels in the product's
        brick-major order (kfb_device_ptr(ctx, 0) of a whole-volume context), float4 model
        maps of level 0, camera pose.  The caller has synchronised the producer's stream."""
        p = _f(pose12)
        lib().ref_kinfu_sync_from(self.h, _vp(int(vol_packed_dev)), _vp(int(vmap4_dev)), _vp(int(nmap4_dev)), _p(p), int(frame_count))


def event_tic():
    lib().ref_event_ms(0)


def event_toc_ms():
    return lib().ref_event_ms(1)


def depth_truncation(depth_mm_filtered, max_dist=5.0):
    d = np.array(depth_mm_filtered, dtype=np.float32, copy=True)
    lib().ref_depth_truncation(_p(d), d.shape[1], d.shape[0], max_dist)
    return d


def vertex_normal(depth_m, K):
    d = _f(depth_m)
    v = np.empty((K.height, K.width, 3), np.float32)
    n = np.empty((K.height, K.width, 3), np.float32)
    lib().ref_vertex_normal(_p(d), K.width, K.height, K.fx, K.fy, K.cx, K.cy, _p(v), _p(n))
    return v, n


def resize_maps(vbig, nbig):
    vb, nb = _f(vbig), _f(nbig)
    h, w, _ = vb.shape
    vs = np.empty((h >> 1, w >> 1, 3), np.float32)
    ns = np.empty((h >> 1, w >> 1, 3), np.float32)
    lib().ref_resize_maps(_p(vb), _p(nb), w, h, _p(vs), _p(ns))
    return vs, ns


def render(vmap, nmap, eye, phong=True):
    v, n, e = _f(vmap), _f(nmap), _f(eye)
    h, w, _ = v.shape
    out = np.zeros((h, w, 3), np.uint8)
    lib().ref_render(_p(v), _p(n), w, h, _p(e), int(phong), _p(out))
    return out


def rigid_icp(cur_v, cur_n, pre_v, pre_n, K, pose12, dist=0.015, sine=0.5, reps=1):
    cv, cn, pv, pn, p = _f(cur_v), _f(cur_n), _f(pre_v), _f(pre_n), _f(pose12)
    A = np.zeros((6, 6), np.float64)
    b = np.zeros(6, np.float64)
    ms = lib().ref_rigid_icp(_p(cv), _p(cn), _p(pv), _p(pn), K.width, K.height, K.fx, K.fy, K.cx, K.cy, _p(p),
                             dist, sine, _p(A), _p(b), reps)
    return A, b, ms


def sums27_from_Ab(A, b):
    out = np.zeros(27, np.float64)
    s = 0
    for i in range(6):
        for j in range(i, 7):
            out[s] = b[i] if j == 6 else A[i, j]
            s += 1
    return out
