"""ctypes binding of libodo_b200.so (include/odo_b200.h) for tests and bench.py.

This is plumbing: every function here forwards to the C-ABI, which is the only way to reach the
CUDA kernels.  There is NO CPU fallback: if the library is missing or no CUDA device is present the
calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libodo_b200.so")
MAX_LEVELS = 8


class Intrinsics(C.Structure):
    _fields_ = [("fx", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class LMParams(C.Structure):
    _fields_ = [("lam", C.c_float), ("precision", C.c_float), ("max_iters", C.c_int * MAX_LEVELS),
                ("robust", C.c_int), ("huber_delta", C.c_float), ("K", Intrinsics)]


class LMStats(C.Structure):
    _fields_ = [("iters", C.c_int * MAX_LEVELS), ("n_res", C.c_int * MAX_LEVELS),
                ("err", C.c_float * MAX_LEVELS), ("status", C.c_int)]


class DepthParams(C.Structure):
    _fields_ = [("grad_th", C.c_float), ("ssd_th", C.c_float), ("photo_th", C.c_float),
                ("min_depth", C.c_float), ("max_depth", C.c_float), ("lam", C.c_float),
                ("huber_delta", C.c_float), ("precision", C.c_float), ("max_iters", C.c_int),
                ("boundary", C.c_int), ("max_residuals", C.c_int), ("baseline", C.c_float),
                ("fx", C.c_float)]


class DepthStats(C.Structure):
    _fields_ = [("n_selected", C.c_int), ("n_matched", C.c_int), ("n_valid", C.c_int),
                ("iters", C.c_int), ("cost", C.c_float), ("status", C.c_int)]


EXPORTS = [
    "odo_version", "odo_ctx_create", "odo_ctx_destroy", "odo_ctx_sync", "odo_last_error",
    "odo_ctx_launch_count", "odo_default_intrinsics", "odo_default_lm_params",
    "odo_default_depth_params", "odo_image_pyramid_build", "odo_depth_pyramid_build",
    "odo_image_pyramid_build_dev", "odo_depth_pyramid_build_dev",
    "odo_pyr_levels", "odo_pyr_level_size", "odo_pyr_download", "odo_pyr_release",
    "odo_compute_depth", "odo_select_pixels", "odo_disparity_search", "odo_depth_optimize",
    "odo_lm_solve", "odo_lm_linearize", "odo_tracker_create", "odo_tracker_destroy",
    "odo_tracker_step", "odo_tracker_step_u8", "odo_tracker_step_dev", "odo_tracker_stage_ms",
    "odo_tracker_enqueue", "odo_tracker_enqueue_u8", "odo_tracker_enqueue_dev", "odo_tracker_collect",
    "odo_tracker_step_dev_u8", "odo_tracker_enqueue_dev_u8",
    "odo_camera_create", "odo_camera_destroy", "odo_camera_configure", "odo_camera_set_maps", "odo_camera_levels",
    "odo_camera_intrinsics", "odo_camera_raw", "odo_camera_lm_intrinsics", "odo_camera_map_size",
    "odo_camera_maps_download", "odo_camera_undistort_rectify", "odo_camera_undistort_rectify_dev",
    "odo_camera_undistort_rectify_batch",
    "odo_tracker_set_rectify", "odo_tracker_enqueue_raw_u8", "odo_tracker_step_raw_u8",
]
PIX_U8, PIX_F32 = 0, 5
STAGES = ["image_pyramid", "lm_solve", "blur_select", "disparity", "depth_refine", "depth_pyr_promote_prepare"]

_lib = None


def lib():
    """Load libodo_b200.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python odometry_b200/build.py` "
                               "(there is no CPU fallback for the product path)")
        L = C.CDLL(LIB_PATH)
        L.odo_version.restype = C.c_char_p
        L.odo_last_error.restype = C.c_char_p
        L.odo_last_error.argtypes = [C.c_void_p]
        L.odo_ctx_launch_count.restype = C.c_int64
        L.odo_ctx_launch_count.argtypes = [C.c_void_p]
        L.odo_ctx_create.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.odo_ctx_destroy.argtypes = [C.c_void_p]
        L.odo_ctx_sync.argtypes = [C.c_void_p]
        L.odo_pyr_release.argtypes = [C.c_void_p, C.c_void_p]
        L.odo_pyr_levels.argtypes = [C.c_void_p]
        L.odo_tracker_destroy.argtypes = [C.c_void_p]
        dp = C.POINTER(C.c_double)
        L.odo_camera_create.argtypes = [C.c_int, dp, dp, C.c_double, C.c_double, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.odo_camera_destroy.argtypes = [C.c_void_p]
        L.odo_camera_configure.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_int, C.c_int]
        L.odo_camera_set_maps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, dp]
        L.odo_camera_levels.argtypes = [C.c_void_p]
        L.odo_camera_intrinsics.argtypes = [C.c_void_p, C.c_int, dp]
        L.odo_camera_raw.argtypes = [C.c_void_p, dp, dp, dp, C.POINTER(C.c_int)]
        L.odo_camera_lm_intrinsics.argtypes = [C.c_void_p, C.POINTER(Intrinsics)]
        L.odo_camera_map_size.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.odo_camera_maps_download.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.odo_camera_undistort_rectify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                   C.c_void_p, C.c_int, C.c_double]
        L.odo_camera_undistort_rectify_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                       C.c_int64, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_double]
        L.odo_camera_undistort_rectify_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                                         C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.odo_tracker_set_rectify.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def _colmajor(T):
    return _f32(np.asarray(T, dtype=np.float32).T).ravel()


def default_lm_params(**kw) -> LMParams:
    p = LMParams()
    lib().odo_default_lm_params(C.byref(p))
    for k, v in kw.items():
        if k == "max_iters":
            for i in range(MAX_LEVELS):
                p.max_iters[i] = v[i] if i < len(v) else 0
        else:
            setattr(p, k, v)
    return p


def default_depth_params(**kw) -> DepthParams:
    p = DepthParams()
    lib().odo_default_depth_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class OdoError(RuntimeError):
    pass


class Pyramid:
    def __init__(self, ctx: "Context", handle):
        self.ctx, self.h = ctx, handle
        ctx._children.add(self)

    @property
    def levels(self):
        return lib().odo_pyr_levels(self.h)

    def level_size(self, l):
        w, h = C.c_int(), C.c_int()
        if lib().odo_pyr_level_size(self.h, l, C.byref(w), C.byref(h)) != 0:
            raise OdoError("bad level")
        return w.value, h.value

    def download(self, l):
        w, h = self.level_size(l)
        out = np.empty((h, w), np.float32)
        if lib().odo_pyr_download(self.ctx.h, self.h, l, _p(out)) != 0:
            raise OdoError(self.ctx.last_error())
        return out

    def release(self):
        if self.h and self.ctx.h:
            lib().odo_pyr_release(self.ctx.h, self.h)
        self.h = None

    close = release

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


class Context:
    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        rc = lib().odo_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0 or not h:
            raise OdoError("odo_ctx_create failed (no CUDA device?) -- there is no CPU fallback")
        self.h = h
        self._children = weakref.WeakSet()   # pyramids / trackers must die before their context

    def last_error(self):
        return lib().odo_last_error(self.h).decode()

    def sync(self):
        if lib().odo_ctx_sync(self.h) != 0:
            raise OdoError(self.last_error())

    def launch_count(self):
        return int(lib().odo_ctx_launch_count(self.h))

    def close(self):
        if self.h:
            for ch in list(self._children):
                ch.close()
            lib().odo_ctx_destroy(self.h)
            self.h = None

    # ---- pyramids
    def image_pyramid(self, img, levels=4, smooth=True) -> Pyramid:
        img = _f32(img)
        hgt, wid = img.shape
        out = C.c_void_p()
        if lib().odo_image_pyramid_build(self.h, _p(img), wid, hgt, levels, int(bool(smooth)), C.byref(out)) != 0:
            raise OdoError(self.last_error())
        return Pyramid(self, out)

    def depth_pyramid(self, dep, levels=4, smooth=False) -> Pyramid:
        dep = _f32(dep)
        hgt, wid = dep.shape
        out = C.c_void_p()
        if lib().odo_depth_pyramid_build(self.h, _p(dep), wid, hgt, levels, int(bool(smooth)), C.byref(out)) != 0:
            raise OdoError(self.last_error())
        return Pyramid(self, out)

    def image_pyramid_dev(self, dev_ptr, wid, hgt, levels=4, smooth=True) -> Pyramid:
        """Image already on the device (float32 [hgt][wid]); stream ordered, returns without waiting."""
        out = C.c_void_p()
        if lib().odo_image_pyramid_build_dev(self.h, C.c_void_p(dev_ptr), wid, hgt, levels, int(bool(smooth)), C.byref(out)) != 0:
            raise OdoError(self.last_error())
        return Pyramid(self, out)

    def depth_pyramid_dev(self, dev_ptr, wid, hgt, levels=4, smooth=False) -> Pyramid:
        out = C.c_void_p()
        if lib().odo_depth_pyramid_build_dev(self.h, C.c_void_p(dev_ptr), wid, hgt, levels, int(bool(smooth)), C.byref(out)) != 0:
            raise OdoError(self.last_error())
        return Pyramid(self, out)

    # ---- LM
    def lm_solve(self, pyr1: Pyramid, dep1: Pyramid, pyr2: Pyramid, params: LMParams, T_init=None):
        T0 = _colmajor(np.eye(4, dtype=np.float32) if T_init is None else T_init)
        Tout = np.empty(16, np.float32)
        st = LMStats()
        rc = lib().odo_lm_solve(self.h, pyr1.h, dep1.h, pyr2.h, C.byref(params), _p(T0), _p(Tout), C.byref(st))
        return rc, Tout.reshape(4, 4).T.copy(), st

    def lm_linearize(self, pyr1, dep1, pyr2, level, T, robust=1, huber_delta=28.0):
        Tc = _colmajor(T)
        H = np.zeros(21, np.float64)
        g = np.zeros(6, np.float64)
        E, N, sigma = C.c_double(), C.c_int(), C.c_float()
        rc = lib().odo_lm_linearize(self.h, pyr1.h, dep1.h, pyr2.h, level, _p(Tc), robust, C.c_float(huber_delta),
                                    None, _p(H, C.c_double), _p(g, C.c_double), C.byref(E), C.byref(N), C.byref(sigma))
        return dict(rc=rc, H=H, g=g, E=E.value, N=N.value, sigma=sigma.value)

    # ---- depth
    def select_pixels(self, L, R, p: DepthParams):
        L, R = _f32(L), _f32(R)
        h, w = L.shape
        bl, br = np.empty_like(L), np.empty_like(R)
        val = np.zeros((h, w), np.uint8)
        rc = lib().odo_select_pixels(self.h, _p(L), _p(R), w, h, C.byref(p), _p(bl), _p(br), _p(val, C.c_uint8))
        return rc, bl, br, val

    def disparity_search(self, bl, br, val, p: DepthParams, want_ssd=True):
        """want_ssd=False: min_ssd is not requested (returned as zeros), which lets the library use its
        exact integer path when both images are multiples of 1/16."""
        bl, br = _f32(bl), _f32(br)
        h, w = bl.shape
        val = np.ascontiguousarray(val, np.uint8)
        disp = np.zeros((h, w), np.float32)
        dep = np.zeros((h, w), np.float32)
        match = np.zeros((h, w), np.int32)
        ssd = np.zeros((h, w), np.float32)
        rc = lib().odo_disparity_search(self.h, _p(bl), _p(br), w, h, C.byref(p), _p(val, C.c_uint8), _p(disp),
                                        _p(dep), _p(match, C.c_int32), _p(ssd) if want_ssd else None)
        return rc, disp, dep, match, ssd

    def depth_optimize(self, L, R, dep, val, p: DepthParams):
        L, R = _f32(L), _f32(R)
        h, w = L.shape
        dep = _f32(dep).copy()
        val = np.ascontiguousarray(val, np.uint8).copy()
        st = DepthStats()
        rc = lib().odo_depth_optimize(self.h, _p(L), _p(R), w, h, C.byref(p), _p(dep), _p(val, C.c_uint8), C.byref(st))
        return rc, dep, val, st

    def compute_depth(self, L, R, p: DepthParams):
        L, R = _f32(L), _f32(R)
        h, w = L.shape
        val = np.zeros((h, w), np.uint8)
        disp = np.zeros((h, w), np.float32)
        dep = np.zeros((h, w), np.float32)
        st = DepthStats()
        rc = lib().odo_compute_depth(self.h, _p(L), _p(R), w, h, C.byref(p), _p(val, C.c_uint8), _p(disp), _p(dep),
                                     C.byref(st))
        return rc, val, disp, dep, st


class Camera:
    """CameraPyramid of the reference (include/camera.h:15-118) over odo_camera_*: raw parameters, ConfigureCamera
    (per-level intrinsics + device-resident remap table) and UndistortRectify."""

    def __init__(self, ctx: Context, levels, K_raw, dist, sensor=(0.0, 0.0), resolution=(0, 0)):
        self.ctx = ctx
        K = np.ascontiguousarray(K_raw, np.float64).reshape(9)
        D = np.ascontiguousarray(dist, np.float64).reshape(4)
        h = C.c_void_p()
        rc = lib().odo_camera_create(levels, _p(K, C.c_double), _p(D, C.c_double), float(sensor[0]), float(sensor[1]),
                                     int(resolution[0]), int(resolution[1]), C.byref(h))
        if rc != 0:
            raise OdoError("odo_camera_create failed")
        self.h = h
        ctx._children.add(self)

    def configure(self, R, P34, new_size):
        """ConfigureCamera(rectify_rotation, new_intrinsic34, new_size), src/camera.cpp:40-70; new_size = (w, h)."""
        R = np.ascontiguousarray(R, np.float64).reshape(9)
        P = np.ascontiguousarray(P34, np.float64).reshape(12)
        rc = lib().odo_camera_configure(self.ctx.h, self.h, _p(R, C.c_double), _p(P, C.c_double), new_size[0], new_size[1])
        if rc != 0:
            raise OdoError(self.ctx.last_error())

    def set_maps(self, mapx, mapy, P34=None):
        mapx, mapy = _f32(mapx), _f32(mapy)
        hh, ww = mapx.shape
        P = None if P34 is None else np.ascontiguousarray(P34, np.float64).reshape(12)
        rc = lib().odo_camera_set_maps(self.ctx.h, self.h, mapx.ctypes.data, mapy.ctypes.data, ww, hh,
                                       None if P is None else _p(P, C.c_double))
        if rc != 0:
            raise OdoError(self.ctx.last_error())

    def intrinsics(self, level):
        out = np.zeros(5, np.float64)
        if lib().odo_camera_intrinsics(self.h, level, _p(out, C.c_double)) != 0:
            raise OdoError("camera not configured or level out of range")
        return out

    def lm_intrinsics(self) -> Intrinsics:
        k = Intrinsics()
        if lib().odo_camera_lm_intrinsics(self.h, C.byref(k)) != 0:
            raise OdoError("camera not configured")
        return k

    def map_size(self):
        w, h = C.c_int(), C.c_int()
        if lib().odo_camera_map_size(self.h, C.byref(w), C.byref(h)) != 0:
            raise OdoError("camera not configured")
        return w.value, h.value

    def maps(self):
        w, h = self.map_size()
        mx, my = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
        if lib().odo_camera_maps_download(self.ctx.h, self.h, mx.ctypes.data, my.ctypes.data) != 0:
            raise OdoError(self.ctx.last_error())
        return mx, my

    def undistort_rectify(self, src, out_dtype=None, border_value=0.0):
        """UndistortRectify(src_raw, dst, INTER_LINEAR, BORDER_CONSTANT, Scalar(border_value)), src/camera.cpp:72-83."""
        src = np.ascontiguousarray(src)
        if src.dtype not in (np.uint8, np.float32):
            raise TypeError("u8 or float32 frames")
        out_dtype = np.dtype(out_dtype or src.dtype)
        w, h = self.map_size()
        dst = np.empty((h, w), out_dtype)
        code = {np.dtype(np.uint8): PIX_U8, np.dtype(np.float32): PIX_F32}
        rc = lib().odo_camera_undistort_rectify(self.ctx.h, self.h, src.ctypes.data, code[src.dtype], src.shape[1],
                                                src.shape[0], dst.ctypes.data, code[out_dtype], float(border_value))
        if rc != 0:
            raise OdoError(self.ctx.last_error())
        return dst

    def undistort_rectify_batch_ptrs(self, src_ptrs, src_u8, sw, sh, dst_ptrs, dst_u8, border_value=0.0):
        """n host images (raw addresses, ideally pinned) -> n host images: one launch, one wait."""
        n = len(src_ptrs)
        arr_t = C.c_void_p * n
        rc = lib().odo_camera_undistort_rectify_batch(self.ctx.h, self.h, arr_t(*src_ptrs), PIX_U8 if src_u8 else PIX_F32, sw, sh,
                                                      arr_t(*dst_ptrs), PIX_U8 if dst_u8 else PIX_F32, n, float(border_value))
        if rc != 0:
            raise OdoError(self.ctx.last_error())

    def undistort_rectify_dev(self, src_ptr, src_u8, sw, sh, src_stride, dst_ptr, dst_u8, dst_stride, batch, border_value=0.0):
        rc = lib().odo_camera_undistort_rectify_dev(self.ctx.h, self.h, C.c_void_p(src_ptr), PIX_U8 if src_u8 else PIX_F32,
                                                    sw, sh, src_stride, C.c_void_p(dst_ptr), PIX_U8 if dst_u8 else PIX_F32,
                                                    dst_stride, batch, float(border_value))
        if rc != 0:
            raise OdoError(self.ctx.last_error())

    def close(self):
        if self.h:
            lib().odo_camera_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tracker:
    """Batched tracking loop (odo_tracker_*): n_seq sequences, one frame per sequence per step."""

    def __init__(self, ctx: Context, w, h, levels=4, n_seq=1, lm: LMParams | None = None,
                 dp: DepthParams | None = None):
        self.ctx, self.w, self.hgt, self.n = ctx, w, h, n_seq
        self._lm = lm or default_lm_params()
        self._dp = dp or default_depth_params()
        t = C.c_void_p()
        rc = lib().odo_tracker_create(ctx.h, w, h, levels, n_seq, C.byref(self._lm), C.byref(self._dp), C.byref(t))
        if rc != 0:
            raise OdoError(ctx.last_error())
        self.h = t
        ctx._children.add(self)
        self.pose_abs = np.zeros((n_seq, 16), np.float32)
        self.pose_kf = np.zeros((n_seq, 16), np.float32)
        self.is_kf = np.zeros(n_seq, np.int32)
        self.status = np.zeros(n_seq, np.int32)

    def _outs(self):
        return (_p(self.pose_abs), _p(self.pose_kf), _p(self.is_kf, C.c_int), _p(self.status, C.c_int))

    def _result(self, rc):
        pa = self.pose_abs.reshape(self.n, 4, 4).transpose(0, 2, 1).copy()
        pk = self.pose_kf.reshape(self.n, 4, 4).transpose(0, 2, 1).copy()
        return rc, pa, pk, self.is_kf.copy(), self.status.copy()

    def step_ptrs(self, left_ptrs, right_ptrs, u8=False):
        """left_ptrs/right_ptrs: sequences of n_seq raw host addresses (e.g. pinned torch tensors)."""
        arr_t = C.c_void_p * self.n
        la, ra = arr_t(*left_ptrs), arr_t(*right_ptrs)
        fn = lib().odo_tracker_step_u8 if u8 else lib().odo_tracker_step
        rc = fn(self.h, la, ra, *self._outs())
        return self._result(rc)

    def enqueue_ptrs(self, left_ptrs, right_ptrs, u8=False):
        """Pipelined ingest: start the copies + kernels of one frame and return at once (<= 2 in flight)."""
        arr_t = C.c_void_p * self.n
        la, ra = arr_t(*left_ptrs), arr_t(*right_ptrs)
        fn = lib().odo_tracker_enqueue_u8 if u8 else lib().odo_tracker_enqueue
        if fn(self.h, la, ra) != 0:
            raise OdoError(self.ctx.last_error())

    def pointer_array(self, ptrs):
        """a reusable C array of n_seq host addresses for enqueue_arrays (what a C++ caller holds anyway)"""
        return (C.c_void_p * self.n)(*ptrs)

    def enqueue_arrays(self, left_arr, right_arr, u8=False):
        """enqueue_ptrs with pointer arrays built ahead of time (no per-frame Python work)"""
        fn = lib().odo_tracker_enqueue_u8 if u8 else lib().odo_tracker_enqueue
        if fn(self.h, left_arr, right_arr) != 0:
            raise OdoError(self.ctx.last_error())

    def set_rectify(self, cam_left: "Camera", cam_right: "Camera", raw_w, raw_h):
        """Raw-camera ingest: later frames go through step_raw_u8 / enqueue_raw_ptrs as unrectified u8 images."""
        self._cams = (cam_left, cam_right)
        if lib().odo_tracker_set_rectify(self.h, cam_left.h, cam_right.h, raw_w, raw_h) != 0:
            raise OdoError(self.ctx.last_error())

    def enqueue_raw_ptrs(self, left_ptrs, right_ptrs):
        arr_t = C.c_void_p * self.n
        if lib().odo_tracker_enqueue_raw_u8(self.h, arr_t(*left_ptrs), arr_t(*right_ptrs)) != 0:
            raise OdoError(self.ctx.last_error())

    def step_raw_u8(self, lefts, rights):
        ls = [np.ascontiguousarray(a, dtype=np.uint8) for a in lefts]
        rs = [np.ascontiguousarray(a, dtype=np.uint8) for a in rights]
        arr_t = C.c_void_p * self.n
        rc = lib().odo_tracker_step_raw_u8(self.h, arr_t(*[a.ctypes.data for a in ls]), arr_t(*[a.ctypes.data for a in rs]),
                                           *self._outs())
        return self._result(rc)

    def enqueue_dev(self, left_dev_ptr, right_dev_ptr, u8=False):
        """frames resident on the device: [n_seq][h][w] float32 per eye, or uint8 with u8=True"""
        fn = lib().odo_tracker_enqueue_dev_u8 if u8 else lib().odo_tracker_enqueue_dev
        if fn(self.h, C.c_void_p(left_dev_ptr), C.c_void_p(right_dev_ptr)) != 0:
            raise OdoError(self.ctx.last_error())

    def collect(self):
        """Results of the oldest frame in flight."""
        rc = lib().odo_tracker_collect(self.h, *self._outs())
        return self._result(rc)

    def step(self, lefts, rights):
        """lefts/rights: lists of n_seq numpy images (float32, or uint8 for the fused-convert path)."""
        u8 = lefts[0].dtype == np.uint8
        dt = np.uint8 if u8 else np.float32
        ls = [np.ascontiguousarray(a, dtype=dt) for a in lefts]
        rs = [np.ascontiguousarray(a, dtype=dt) for a in rights]
        return self.step_ptrs([a.ctypes.data for a in ls], [a.ctypes.data for a in rs], u8=u8)

    def step_dev(self, left_dev_ptr, right_dev_ptr, u8=False):
        fn = lib().odo_tracker_step_dev_u8 if u8 else lib().odo_tracker_step_dev
        rc = fn(self.h, C.c_void_p(left_dev_ptr), C.c_void_p(right_dev_ptr), *self._outs())
        return self._result(rc)

    def stage_ms(self):
        out = (C.c_float * len(STAGES))()
        lib().odo_tracker_stage_ms(self.h, out)
        return dict(zip(STAGES, [float(v) for v in out]))

    def close(self):
        if self.h and self.ctx.h:
            lib().odo_tracker_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
