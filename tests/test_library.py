"""The C-ABI library builds, loads and exports every symbol include/odo_b200.h declares -- CPU only."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libpath():
    from odometry_b200 import build
    return build.build()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "odo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(odo_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(libpath):
    lib = ctypes.CDLL(libpath)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/odo_b200.h but not exported"


def test_binding_lists_the_same_symbols(libpath):
    from odometry_b200 import capi
    assert sorted(capi.EXPORTS) == declared_symbols()


def test_defaults_match_the_reference_constants(libpath):
    from odometry_b200 import capi
    lm = capi.default_lm_params()
    assert (lm.lam, lm.precision, lm.robust, lm.huber_delta) == pytest.approx((0.01, 0.995, 1, 28.0))
    assert list(lm.max_iters)[:4] == [10, 20, 30, 30]
    assert (lm.K.fx, lm.K.cx, lm.K.cy) == pytest.approx((718.856, 607.1928, 185.2157), rel=1e-6)
    dp = capi.default_depth_params()
    assert (dp.grad_th, dp.ssd_th, dp.photo_th, dp.min_depth, dp.max_depth) == pytest.approx((8, 900, 15, 0.1, 30))
    assert (dp.max_iters, dp.boundary, dp.max_residuals) == (50, 4, 80000)
    assert dp.baseline == pytest.approx(386.1448 / 718.856, rel=1e-6)


def test_no_cpu_fallback_without_a_device(libpath):
    """On a box without a GPU the product must fail loudly, never compute on the CPU."""
    import torch
    from odometry_b200 import capi
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.OdoError):
        capi.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "odometry_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "liboracle" not in txt and '"oracle.h"' not in txt, f
