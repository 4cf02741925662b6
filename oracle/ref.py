"""ctypes binding of oracle/_ref/libodoref.so: the REFERENCE'S OWN hot-path sources compiled from
/root/reference against the stand-in headers of oracle/ref_shim (see oracle/ref_api.cpp, oracle/Makefile).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__ and bench.py's cpu_baseline /
--impl reference legs.  The library is built in the container that has /root/reference and travels to
the GPU box as a prebuilt file; available() is False where neither the sources nor the file exist.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import (DepthParams, DepthStats, LMParams, SUM_WIDE, _colmajor, _f32, _p, depth_params, lm_params)  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libodoref.so")
REFERENCE_ROOT = os.environ.get("ODO_REFERENCE_ROOT", "/root/reference")

_lib = None


def sources_present() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "src", "lm_optimizer.cpp"))


def build(force: bool = False) -> str | None:
    """make -C oracle _ref (needs the reference sources; a prebuilt library is used as is)."""
    if sources_present():
        subprocess.check_call(["make", "-C", _HERE, "REF=" + REFERENCE_ROOT, "_ref"] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available() -> bool:
    return os.path.exists(_LIB_PATH) or sources_present()


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_describe.restype = C.c_char_p
        _lib.ref_tracker_create.restype = C.c_void_p
        _lib.ref_tracker_destroy.argtypes = [C.c_void_p]
    return _lib


def describe() -> str:
    return lib().ref_describe().decode()


def _pyr_floats(w, h, levels):
    return sum((w >> l) * (h >> l) for l in range(levels))


def image_pyramid(img, levels, smooth=True):
    img = _f32(img)
    h, w = img.shape
    buf = np.empty(_pyr_floats(w, h, levels), np.float32)
    lib().ref_image_pyramid(_p(img), w, h, levels, int(bool(smooth)), _p(buf))
    return buf


def depth_pyramid(dep, levels, smooth=False):
    dep = _f32(dep)
    h, w = dep.shape
    buf = np.empty(_pyr_floats(w, h, levels), np.float32)
    lib().ref_depth_pyramid(_p(dep), w, h, levels, int(bool(smooth)), _p(buf))
    return buf


def lm_linearize(I1, I2, D1, level, T, robust=1, huber_delta=28.0, sum_mode=SUM_WIDE):
    I1, I2, D1 = _f32(I1), _f32(I2), _f32(D1)
    h, w = I1.shape
    H = np.zeros(21, np.float64)
    g = np.zeros(6, np.float64)
    E = C.c_double()
    N = C.c_int()
    res = np.zeros(w * h, np.float32)
    jac = np.zeros((w * h, 6), np.float32)
    wgt = np.zeros(w * h, np.float32)
    rc = lib().ref_lm_linearize(_p(I1), _p(I2), _p(D1), w, h, level, _p(_colmajor(T)), robust, C.c_float(huber_delta),
                                int(sum_mode), _p(H, C.c_double), _p(g, C.c_double), C.byref(E), C.byref(N),
                                _p(res), _p(jac), _p(wgt))
    n = N.value
    return dict(rc=rc, H=H, g=g, E=E.value, N=n, res=res[:n], jac=jac[:n], wgt=wgt[:n])


def lm_solve(img1, dep1, img2, levels, params: LMParams, T_init=None, smooth1=True, smooth2=True):
    img1, dep1, img2 = _f32(img1), _f32(dep1), _f32(img2)
    h, w = img1.shape
    T0 = _colmajor(np.eye(4, dtype=np.float32) if T_init is None else T_init)
    Tout = np.empty(16, np.float32)
    rc = lib().ref_lm_solve(_p(img1), int(smooth1), _p(dep1), _p(img2), int(smooth2), w, h, levels, C.byref(params),
                            _p(T0), _p(Tout))
    return rc, Tout.reshape(4, 4).T.copy()


def disparity(L, R, p: DepthParams):
    L, R = _f32(L), _f32(R)
    h, w = L.shape
    val = np.zeros((h, w), np.uint8)
    disp = np.zeros((h, w), np.float32)
    dep = np.zeros((h, w), np.float32)
    rc = lib().ref_disparity(_p(L), _p(R), w, h, C.byref(p), _p(val, C.c_uint8), _p(disp), _p(dep))
    return rc, val, disp, dep


def depth_optimize(L, R, dep, val, p: DepthParams):
    L, R = _f32(L), _f32(R)
    h, w = L.shape
    dep = _f32(dep).copy()
    val = np.ascontiguousarray(val, np.uint8).copy()
    st = DepthStats()
    rc = lib().ref_depth_optimize(_p(L), _p(R), w, h, C.byref(p), _p(dep), _p(val, C.c_uint8), C.byref(st))
    return rc, dep, val, st


def compute_depth(L, R, p: DepthParams):
    L, R = _f32(L), _f32(R)
    h, w = L.shape
    val = np.zeros((h, w), np.uint8)
    disp = np.zeros((h, w), np.float32)
    dep = np.zeros((h, w), np.float32)
    st = DepthStats()
    rc = lib().ref_compute_depth(_p(L), _p(R), w, h, C.byref(p), _p(val, C.c_uint8), _p(disp), _p(dep), C.byref(st))
    return rc, val, disp, dep, st


class Tracker:
    """run_odometry_kitti_offline.cpp:95-271 over the reference's classes."""

    def __init__(self, w, h, levels=4, lm: LMParams | None = None, dp: DepthParams | None = None):
        self._lm = lm or lm_params()
        self._dp = dp or depth_params()
        self._t = C.c_void_p(lib().ref_tracker_create(w, h, levels, C.byref(self._lm), C.byref(self._dp)))

    def step(self, L, R):
        L, R = _f32(L), _f32(R)
        pa = np.empty(16, np.float32)
        pk = np.empty(16, np.float32)
        kf = C.c_int()
        mag = C.c_float()
        rc = lib().ref_tracker_step(self._t, _p(L), _p(R), _p(pa), _p(pk), C.byref(kf), C.byref(mag))
        return rc, pa.reshape(4, 4).T.copy(), pk.reshape(4, 4).T.copy(), bool(kf.value), mag.value

    def close(self):
        if self._t:
            lib().ref_tracker_destroy(self._t)
            self._t = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
