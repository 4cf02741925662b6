"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (odometry_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

SUM_FLOAT_SEQ = 0
SUM_WIDE = 1
SUM_PACKET4 = 2     # oracle/_ref only: other legal fp32 associations of the N-term products
SUM_PAIRWISE = 3    # (tools/association_noise.py)


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (g++ only)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.exists(os.path.join(_HERE, "Makefile")) and _stale():
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB_PATH


def _stale() -> bool:
    """the library is older than one of its sources (make decides; this only avoids spawning it on every import)"""
    t = os.path.getmtime(_LIB_PATH)
    return any(f.endswith((".cpp", ".h")) and os.path.getmtime(os.path.join(_HERE, f)) > t for f in os.listdir(_HERE))


class Intrinsics(C.Structure):
    _fields_ = [("fx", C.c_float), ("cx", C.c_float), ("cy", C.c_float)]


class LMParams(C.Structure):
    _fields_ = [("lam", C.c_float), ("precision", C.c_float), ("max_iters", C.c_int * 8),
                ("robust", C.c_int), ("huber_delta", C.c_float), ("sum_mode", C.c_int),
                ("K", Intrinsics)]


class LMStats(C.Structure):
    _fields_ = [("iters", C.c_int * 8), ("n_res", C.c_int * 8), ("err", C.c_float * 8),
                ("status", C.c_int)]


class DepthParams(C.Structure):
    _fields_ = [("grad_th", C.c_float), ("ssd_th", C.c_float), ("photo_th", C.c_float),
                ("min_depth", C.c_float), ("max_depth", C.c_float), ("lam", C.c_float),
                ("huber_delta", C.c_float), ("precision", C.c_float), ("max_iters", C.c_int),
                ("boundary", C.c_int), ("max_residuals", C.c_int), ("baseline", C.c_float),
                ("sum_mode", C.c_int), ("fx", C.c_float)]


class DepthStats(C.Structure):
    _fields_ = [("n_selected", C.c_int), ("n_matched", C.c_int), ("n_valid", C.c_int),
                ("iters", C.c_int), ("cost", C.c_float), ("status", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_get_cx_level.restype = C.c_float
        _lib.orc_get_cx_level.argtypes = [C.c_float, C.c_int]
        _lib.orc_pyramid_floats.restype = C.c_int64
        _lib.orc_pyramid_level.restype = C.c_int64
        _lib.orc_motion_magnitude.restype = C.c_float
        _lib.orc_tracker_create.restype = C.c_void_p
        _lib.orc_tracker_destroy.argtypes = [C.c_void_p]
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


def default_intrinsics() -> Intrinsics:
    k = Intrinsics()
    lib().orc_default_intrinsics(C.byref(k))
    return k


def get_cx_level(c: float, level: int) -> float:
    return float(lib().orc_get_cx_level(C.c_float(c), level))


def gaussian_blur3(img):
    img = _f32(img)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_gaussian_blur3(_p(img), w, h, _p(out))
    return out


def pyr_down(img, dsize=None, simd_lanes=4):
    img = _f32(img)
    h, w = img.shape
    dw, dh = dsize if dsize is not None else ((w + 1) // 2, (h + 1) // 2)
    out = np.empty((dh, dw), np.float32)
    rc = lib().orc_pyr_down(_p(img), w, h, _p(out), dw, dh, simd_lanes)
    if rc != 0:
        raise ValueError("pyr_down: bad size")
    return out


def median_blur3(img):
    img = _f32(img)
    h, w = img.shape
    out = np.empty_like(img)
    lib().orc_median_blur3(_p(img), w, h, _p(out))
    return out


def level_shape(w, h, l):
    lw, lh = C.c_int(), C.c_int()
    off = lib().orc_pyramid_level(w, h, l, C.byref(lw), C.byref(lh))
    return int(off), lw.value, lh.value


def split_levels(packed, w, h, levels):
    out = []
    for l in range(levels):
        off, lw, lh = level_shape(w, h, l)
        out.append(packed[off:off + lw * lh].reshape(lh, lw))
    return out


def image_pyramid(img, levels, smooth=True, packed=False):
    img = _f32(img)
    h, w = img.shape
    buf = np.empty(int(lib().orc_pyramid_floats(w, h, levels)), np.float32)
    rc = lib().orc_image_pyramid(_p(img), w, h, levels, int(bool(smooth)), _p(buf))
    if rc != 0:
        raise ValueError("image_pyramid failed")
    return buf if packed else split_levels(buf, w, h, levels)


def depth_pyramid(dep, levels, smooth=False, packed=False):
    dep = _f32(dep)
    h, w = dep.shape
    buf = np.empty(int(lib().orc_pyramid_floats(w, h, levels)), np.float32)
    rc = lib().orc_depth_pyramid(_p(dep), w, h, levels, int(bool(smooth)), _p(buf))
    if rc != 0:
        raise ValueError("depth_pyramid failed")
    return buf if packed else split_levels(buf, w, h, levels)


def se3_exp(twist):
    tw = _f32(twist)
    T = np.empty(16, np.float32)
    lib().orc_se3_exp(_p(tw), _p(T))
    return T.reshape(4, 4).T.copy()  # column-major -> numpy [row, col]


def se3_rewrap(M):
    Mc = _f32(np.asarray(M).T).ravel()
    T = np.empty(16, np.float32)
    lib().orc_se3_rewrap(_p(Mc), _p(T))
    return T.reshape(4, 4).T.copy()


def mat4_mul(A, B):
    a = _f32(np.asarray(A).T).ravel()
    b = _f32(np.asarray(B).T).ravel()
    c = np.empty(16, np.float32)
    lib().orc_mat4_mul(_p(a), _p(b), _p(c))
    return c.reshape(4, 4).T.copy()


def solve6(A, b):
    a = _f32(np.asarray(A).T).ravel()
    bb = _f32(b)
    x = np.empty(6, np.float32)
    lib().orc_solve6(_p(a), _p(bb), _p(x))
    return x


def _colmajor(T):
    return _f32(np.asarray(T, dtype=np.float32).T).ravel()


def lm_params(lam=0.01, precision=0.995, max_iters=(10, 20, 30, 30), robust=1, huber_delta=28.0,
              sum_mode=SUM_WIDE) -> LMParams:
    p = LMParams()
    p.lam = lam
    p.precision = precision
    for i in range(8):
        p.max_iters[i] = max_iters[i] if i < len(max_iters) else 0
    p.robust = robust
    p.huber_delta = huber_delta
    p.sum_mode = sum_mode
    p.K = default_intrinsics()
    return p


def lm_linearize(I1, I2, D1, level, T, robust=1, huber_delta=28.0, sum_mode=SUM_WIDE, details=False):
    """One ComputeResidualJacobianNaive + normal-equation sums. T is a numpy 4x4 [row, col]."""
    I1, I2, D1 = _f32(I1), _f32(I2), _f32(D1)
    h, w = I1.shape
    Tc = _colmajor(T)
    H = np.zeros(21, np.float64)
    g = np.zeros(6, np.float64)
    E = C.c_double()
    N = C.c_int()
    sigma = C.c_float()
    K = default_intrinsics()
    res = jac = wgt = src = dst = None
    args = [None] * 5
    if details:
        res = np.zeros(w * h, np.float32)
        jac = np.zeros((w * h, 6), np.float32)
        wgt = np.zeros(w * h, np.float32)
        src = np.zeros((w * h, 2), np.int32)
        dst = np.zeros((w * h, 2), np.int32)
        args = [_p(res), _p(jac), _p(wgt), _p(src, C.c_int32), _p(dst, C.c_int32)]
    rc = lib().orc_lm_linearize(_p(I1), _p(I2), _p(D1), w, h, level, _p(Tc), robust,
                                C.c_float(huber_delta), sum_mode, C.byref(K),
                                _p(H, C.c_double), _p(g, C.c_double), C.byref(E), C.byref(N),
                                C.byref(sigma), *args)
    out = dict(rc=rc, H=H, g=g, E=E.value, N=N.value, sigma=sigma.value)
    if details:
        n = N.value
        out.update(res=res[:n], jac=jac[:n], wgt=wgt[:n], src=src[:n], dst=dst[:n])
    return out


def lm_solve(pyr1, dep1, pyr2, w, h, levels, params: LMParams, T_init=None):
    """pyr1/dep1/pyr2: packed pyramids. Returns (rc, T 4x4 numpy [row,col], LMStats)."""
    T0 = _colmajor(np.eye(4, dtype=np.float32) if T_init is None else T_init)
    Tout = np.empty(16, np.float32)
    st = LMStats()
    rc = lib().orc_lm_solve(_p(_f32(pyr1)), _p(_f32(dep1)), _p(_f32(pyr2)), w, h, levels,
                            C.byref(params), _p(T0), _p(Tout), C.byref(st))
    return rc, Tout.reshape(4, 4).T.copy(), st


def depth_params(**kw) -> DepthParams:
    p = DepthParams()
    lib().orc_default_depth_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def select_pixels(L, R, p: DepthParams):
    L, R = _f32(L), _f32(R)
    h, w = L.shape
    bl, br = np.empty_like(L), np.empty_like(R)
    val = np.zeros((h, w), np.uint8)
    rc = lib().orc_select_pixels(_p(L), _p(R), w, h, C.byref(p), _p(bl), _p(br), _p(val, C.c_uint8))
    return rc, bl, br, val


def disparity_search(bl, br, val, p: DepthParams):
    bl, br = _f32(bl), _f32(br)
    h, w = bl.shape
    disp = np.zeros((h, w), np.float32)
    dep = np.zeros((h, w), np.float32)
    match = np.full((h, w), -1, np.int32)
    ssd = np.zeros((h, w), np.float32)
    val = np.ascontiguousarray(val, np.uint8)
    rc = lib().orc_disparity_search(_p(bl), _p(br), w, h, C.byref(p), _p(val, C.c_uint8), _p(disp),
                                    _p(dep), _p(match, C.c_int32), _p(ssd))
    return rc, disp, dep, match, ssd


def depth_optimize(L, R, dep, val, p: DepthParams):
    L, R = _f32(L), _f32(R)
    h, w = L.shape
    dep = _f32(dep).copy()
    val = np.ascontiguousarray(val, np.uint8).copy()
    st = DepthStats()
    rc = lib().orc_depth_optimize(_p(L), _p(R), w, h, C.byref(p), _p(dep), _p(val, C.c_uint8), C.byref(st))
    return rc, dep, val, st


def compute_depth(L, R, p: DepthParams):
    L, R = _f32(L), _f32(R)
    h, w = L.shape
    val = np.zeros((h, w), np.uint8)
    disp = np.zeros((h, w), np.float32)
    dep = np.zeros((h, w), np.float32)
    st = DepthStats()
    rc = lib().orc_compute_depth(_p(L), _p(R), w, h, C.byref(p), _p(val, C.c_uint8), _p(disp), _p(dep), C.byref(st))
    return rc, val, disp, dep, st


def camera_intrinsic_pyramid(P34, levels):
    """ConfigureCamera's per-level (fx, fy, cx, cy, f_theta), src/camera.cpp:44-66."""
    P = np.ascontiguousarray(P34, np.float64).reshape(12)
    out = np.empty((levels, 5), np.float64)
    lib().orc_camera_intrinsic_pyramid(_p(P, C.c_double), levels, _p(out, C.c_double))
    return out


def init_undistort_rectify_map(K, dist, R, P34, size):
    """cv::initUndistortRectifyMap(..., CV_32FC1) -> (mapx, mapy); size = (w, h)."""
    w, h = size
    K, dist, R, P = [np.ascontiguousarray(a, np.float64).ravel() for a in (K, dist, R, P34)]
    mx, my = np.empty((h, w), np.float32), np.empty((h, w), np.float32)
    rc = lib().orc_init_undistort_rectify_map(_p(K, C.c_double), _p(dist, C.c_double), _p(R, C.c_double),
                                              _p(P, C.c_double), w, h, _p(mx), _p(my))
    if rc != 0:
        raise ValueError("singular rectification")
    return mx, my


def remap_bilinear(src, mapx, mapy, cval=0):
    """cv::remap(src, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT, cval) for u8 or float32 sources."""
    mapx, mapy = _f32(mapx), _f32(mapy)
    dh, dw = mapx.shape
    sh, sw = src.shape
    if src.dtype == np.uint8:
        src = np.ascontiguousarray(src)
        dst = np.empty((dh, dw), np.uint8)
        lib().orc_remap_bilinear_u8(_p(src, C.c_uint8), sw, sh, _p(mapx), _p(mapy), dw, dh, int(cval), _p(dst, C.c_uint8))
    else:
        src = _f32(src)
        dst = np.empty((dh, dw), np.float32)
        lib().orc_remap_bilinear_f32(_p(src), sw, sh, _p(mapx), _p(mapy), dw, dh, C.c_float(cval), _p(dst))
    return dst


def motion_magnitude(T):
    return float(lib().orc_motion_magnitude(_p(_colmajor(T))))


class Tracker:
    """run_odometry_kitti_offline.cpp loop over the oracle."""

    def __init__(self, w, h, levels=4, lm: LMParams | None = None, dp: DepthParams | None = None, dedup=False):
        self.w, self.h = w, h
        self._lm = lm or lm_params()
        self._dp = dp or depth_params()
        self._t = C.c_void_p(lib().orc_tracker_create(w, h, levels, C.byref(self._lm), C.byref(self._dp), int(dedup)))

    def step(self, L, R):
        L, R = _f32(L), _f32(R)
        pa = np.empty(16, np.float32)
        pk = np.empty(16, np.float32)
        kf = C.c_int()
        mag = C.c_float()
        rc = lib().orc_tracker_step(self._t, _p(L), _p(R), _p(pa), _p(pk), C.byref(kf), C.byref(mag))
        return rc, pa.reshape(4, 4).T.copy(), pk.reshape(4, 4).T.copy(), bool(kf.value), mag.value

    def close(self):
        if self._t:
            lib().orc_tracker_destroy(self._t)
            self._t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
