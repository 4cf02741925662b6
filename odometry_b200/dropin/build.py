"""Build libodometry_dropin.so (the reference's C++ class interfaces over the C-ABI) and the test
driver that uses them like the reference's own callers do.

g++ only: this layer contains no CUDA; it reaches the GPU exclusively through libodo_b200.so.
Without Eigen / OpenCV C++ headers in the image the stand-ins under compat/ are put on the include
path; on a machine that has the real libraries drop `-I compat` and link opencv_core instead.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libodometry_dropin.so")
DRIVER = os.path.join(HERE, "dropin_driver")
SRCS = ["runtime.cpp", "image_pyramid.cpp", "image_processing_global.cpp", "depth_estimate.cpp", "lm_optimizer.cpp",
        "camera.cpp", "stereo_rectify.cpp"]
INCLUDES = ["-I", os.path.join(HERE, "include"), "-I", os.path.join(HERE, "compat"), "-I", os.path.join(ROOT, "include")]
CXXFLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-ffp-contract=off"]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _deps():
    out = []
    for d in ("include", "compat/Eigen", "compat/opencv2", "src"):
        p = os.path.join(HERE, d)
        out += [os.path.join(p, f) for f in os.listdir(p)]
    out.append(os.path.join(ROOT, "include", "odo_b200.h"))
    return out


def build(force: bool = False) -> tuple[str, str]:
    from odometry_b200 import build as core
    core_lib = core.build()
    cxx = os.environ.get("CXX", "g++")
    srcs = [os.path.join(HERE, "src", s) for s in SRCS]
    if force or _stale(LIB, _deps() + [core_lib]):
        cmd = [cxx] + CXXFLAGS + INCLUDES + ["-shared", "-o", LIB] + srcs + \
              ["-L", os.path.dirname(core_lib), "-lodo_b200", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    drv_src = os.path.join(ROOT, "tests", "cpp", "dropin_driver.cpp")
    if force or _stale(DRIVER, [drv_src, LIB] + _deps()):
        cmd = [cxx] + CXXFLAGS + INCLUDES + ["-o", DRIVER, drv_src, "-L", HERE, "-lodometry_dropin",
                                             "-L", os.path.dirname(core_lib), "-lodo_b200",
                                             "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/.."]
        subprocess.check_call(cmd)
    return LIB, DRIVER


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    print(build(force="--force" in sys.argv))
