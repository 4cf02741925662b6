"""Oracle vs golden vectors (cv2 4.13.0) for the undistort + rectify path in front of the tracking loop
(CameraPyramid::ConfigureCamera / UndistortRectify, src/camera.cpp:40-83) -- CPU only.  The fixture was made by
tests/golden/make_golden.py from the reference's own calibration file."""
import os
import sys

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import camera_patterns as cp  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "cv2_camera.npz"))
SIDES = [("left", "K0", "D0", "R1", "P1"), ("right", "K1", "D1", "R2", "P2")]


def sha(a):
    return str(cp.digest(a))


def side_maps(side):
    name, K, D, R, P = side
    return oracle.init_undistort_rectify_map(GOLD[K], GOLD[D], GOLD[R], GOLD[P], cp.RAW_SIZE)


def test_fixture_inputs_are_reproducible():
    assert sha(cp.raw_pattern_u8()) == str(GOLD["src8_sha"])
    assert sha(cp.raw_pattern_f32()) == str(GOLD["srcf_sha"])
    assert sha(np.stack(cp.stress_maps())) == str(GOLD["stress/maps_sha"])


def fixed_point(m):
    """the 1/32-pixel fixed-point coordinate remap() derives from a float map (cvRound(m * 32))"""
    return np.rint(m * np.float32(32)).astype(np.int64)


@pytest.mark.parametrize("side", SIDES, ids=["left", "right"])
def test_rectify_maps_match_cv2_golden(side):
    """Bit-exact against cv2 except where u = fx * xd + u0 cancels to ~0 (one element of the right camera's map,
    |v| = 2e-5 px): there the last float bit depends on FMA contraction inside OpenCV's AVX2 build.  Allowed: at
    most 2 elements, 1 float ulp, < 1e-9 px -- and the fixed-point coordinates remap() uses must be identical,
    which test_remap_matches_cv2_golden then confirms end to end (rectified images bit-exact from OUR maps)."""
    mx, my = side_maps(side)
    name = side[0]
    for m, key in ((mx, "mapx"), (my, "mapy")):
        ref = GOLD[f"{name}/{key}_rows"]
        got = m[::16]
        bad = got.view(np.uint32) != ref.view(np.uint32)
        assert bad.sum() <= 2
        assert np.abs(got.astype(np.float64) - ref.astype(np.float64)).max() < 1e-9
        assert np.array_equal(fixed_point(got), fixed_point(ref))
    if name == "left":
        assert sha(mx) == str(GOLD["left/mapx_sha"]) and sha(my) == str(GOLD["left/mapy_sha"])
    else:
        assert sha(mx) == str(GOLD["right/mapx_sha"])


@pytest.mark.parametrize("side", SIDES, ids=["left", "right"])
@pytest.mark.parametrize("tag", ["u8", "f32"])
@pytest.mark.parametrize("cval", [0, 9])
def test_remap_matches_cv2_golden(side, tag, cval):
    mx, my = side_maps(side)
    src = cp.raw_pattern_u8() if tag == "u8" else cp.raw_pattern_f32()
    d = oracle.remap_bilinear(src, mx, my, cval)
    assert np.array_equal(d[::16], GOLD[f"{side[0]}/remap_{tag}_c{cval}_rows"])
    assert sha(d) == str(GOLD[f"{side[0]}/remap_{tag}_c{cval}_sha"])


@pytest.mark.parametrize("tag", ["u8", "f32"])
def test_remap_stress_maps_match_cv2_golden(tag):
    """maps that leave the raw frame on every side: partial and empty 2x2 footprints take the border value"""
    mx, my = cp.stress_maps()
    src = cp.raw_pattern_u8() if tag == "u8" else cp.raw_pattern_f32()
    d = oracle.remap_bilinear(src, mx, my, 9)
    assert np.array_equal(d[::16], GOLD[f"stress/remap_{tag}_rows"])
    assert sha(d) == str(GOLD[f"stress/remap_{tag}_sha"])


def test_intrinsic_pyramid_follows_configure_camera():
    """src/camera.cpp:44-66: f halves, c <- (c + 0.5) / 2 + 0.5 per level, in double"""
    P = GOLD["P1"]
    k = oracle.camera_intrinsic_pyramid(P, 4)
    fx, fy, cx, cy = P[0, 0], P[1, 1], P[0, 2], P[1, 2]
    for l in range(4):
        assert k[l, 0] == fx and k[l, 1] == fy and k[l, 2] == cx and k[l, 3] == cy and k[l, 4] == 0.0
        fx, fy, cx, cy = fx / 2.0, fy / 2.0, (cx + 0.5) / 2.0 + 0.5, (cy + 0.5) / 2.0 + 0.5
    # the float version of the same recurrence is what the hot path applies per level (GetCxLevel)
    assert abs(oracle.get_cx_level(float(P[0, 2]), 3) - k[3, 2]) < 1e-3


def test_live_cv2_if_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    mx, my = side_maps(SIDES[0])
    src8 = rng.integers(0, 256, (480, 640), dtype=np.uint8)
    srcf = (rng.random((480, 640)) * 255).astype(np.float32)
    for src in (src8, srcf):
        ref = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=3)
        assert np.array_equal(oracle.remap_bilinear(src, mx, my, 3), ref)
    # a different output size and source size (nothing in the algorithm is 640x480 specific)
    mx2 = (rng.random((100, 333), dtype=np.float32) * 90 - 5).astype(np.float32)
    my2 = (rng.random((100, 333), dtype=np.float32) * 70 - 5).astype(np.float32)
    s2 = rng.integers(0, 256, (61, 83), dtype=np.uint8)
    ref = cv2.remap(s2, mx2, my2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    assert np.array_equal(oracle.remap_bilinear(s2, mx2, my2, 0), ref)
