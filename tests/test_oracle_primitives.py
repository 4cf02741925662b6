"""Oracle vs golden vectors (cv2 4.13.0) for the OpenCV primitives on the hot path -- CPU only."""
import os

import numpy as np
import pytest

import oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_primitives.npz"))
CASES = ["u8_157x93", "f32_157x93", "u8_64x48"]


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", CASES)
def test_blur_matches_cv2_golden(name):
    assert bits_equal(oracle.gaussian_blur3(GOLD[f"{name}/img"]), GOLD[f"{name}/blur"])


@pytest.mark.parametrize("name", CASES)
def test_median_matches_cv2_golden(name):
    assert bits_equal(oracle.median_blur3(GOLD[f"{name}/img"]), GOLD[f"{name}/median"])


@pytest.mark.parametrize("name", CASES)
def test_pyrdown_chain_matches_cv2_golden(name):
    cur = GOLD[f"{name}/img"]
    for l in range(1, 4):
        ref = GOLD[f"{name}/down{l}"]
        got = oracle.pyr_down(cur, (ref.shape[1], ref.shape[0]), 4)
        assert bits_equal(got, ref), f"level {l}"
        cur = ref


def test_live_cv2_if_available():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for (w, h) in [(1241, 376), (333, 77)]:
        img = rng.integers(0, 256, (h, w)).astype(np.float32)
        assert bits_equal(oracle.gaussian_blur3(img), cv2.GaussianBlur(img, (3, 3), 0))
        cur = img
        for _ in range(3):
            dw, dh = cur.shape[1] // 2, cur.shape[0] // 2
            ref = cv2.pyrDown(cur, dstsize=(dw, dh))
            assert bits_equal(oracle.pyr_down(cur, (dw, dh), 4), ref)
            cur = ref


def test_image_pyramid_structure():
    """L0 = blur(img), L1 = pyrDown(RAW img) (quirk Q5), Ll = pyrDown(L(l-1)); sizes halve by integer division."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (93, 157)).astype(np.float32)
    lv = oracle.image_pyramid(img, 4, smooth=True)
    assert [a.shape for a in lv] == [(93, 157), (46, 78), (23, 39), (11, 19)]
    assert bits_equal(lv[0], oracle.gaussian_blur3(img))
    assert bits_equal(lv[1], oracle.pyr_down(img, (78, 46)))
    assert bits_equal(lv[2], oracle.pyr_down(lv[1], (39, 23)))
    assert bits_equal(oracle.image_pyramid(img, 2, smooth=False)[0], img)
    # u8-sourced data: levels 0..2 are exact in real arithmetic
    l1_exact = oracle.pyr_down(img.astype(np.float64).astype(np.float32), (78, 46), 0)
    assert bits_equal(lv[1], l1_exact)


def test_depth_pyramid_is_odd_pixel_subsample():
    rng = np.random.default_rng(5)
    dep = rng.random((93, 157)).astype(np.float32)
    lv = oracle.depth_pyramid(dep, 4, smooth=False)
    assert bits_equal(lv[0], dep)
    assert bits_equal(lv[1], dep[1::2, 1::2][:46, :78])
    assert bits_equal(lv[2], lv[1][1::2, 1::2][:23, :39])
    assert bits_equal(lv[3], dep[7::8, 7::8][:11, :19])
    sm = oracle.depth_pyramid(dep, 2, smooth=True)
    assert bits_equal(sm[0], oracle.median_blur3(dep))


def test_get_cx_level_values():
    # SURVEY App. A.3 table
    k = oracle.default_intrinsics()
    want_cx = [607.19281, 304.346405, 152.923203, 77.2116013, 39.3558006, 20.4279003]
    want_cy = [185.215698, 93.3578491, 47.4289246, 24.4644623, 12.9822311, 7.24111557]
    for l in range(6):
        assert oracle.get_cx_level(k.cx, l) == pytest.approx(want_cx[l], rel=1e-7)
        assert oracle.get_cx_level(k.cy, l) == pytest.approx(want_cy[l], rel=1e-7)
