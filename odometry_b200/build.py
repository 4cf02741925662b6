"""Build libodo_b200.so (hand-written CUDA kernels + C-ABI) in-tree with nvcc for sm_100a.

-fmad=false: parity with the reference needs every fp32/fp64 multiply and add rounded separately
(the kernels call __fmaf_rn explicitly where the reference's Eigen code is fused).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libodo_b200.so")
SOURCES = ["pyramid.cu", "lm.cu", "depth.cu", "tracker.cu", "capi.cu", "camera.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-cudart", "static",
]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "include", "odo_b200.h"))
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    if not force and _newer(LIB, srcs + hdrs):
        return LIB
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    objs = [os.path.join(HERE, "build", s.replace(".cu", ".o")) for s in SOURCES]

    def compile_one(i):
        if not force and _newer(objs[i], [srcs[i]] + hdrs):
            return
        extra = os.environ.get("ODO_EXTRA_NVCC", "").split()   # experiment knob (e.g. -DODO_LM_THREADS=256)
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcs[i], "-o", objs[i]]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(compile_one, range(len(SOURCES))))
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
