"""K1/K2 parity: device pyramids vs the oracle (bit-exact) and the cv2 golden vectors. Needs a GPU."""
import os

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "cv2_primitives.npz"))


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", ["u8_157x93", "f32_157x93", "u8_64x48"])
def test_image_pyramid_vs_cv2_golden(ctx, name):
    img = GOLD[f"{name}/img"]
    p = ctx.image_pyramid(img, 3, smooth=True)
    assert bits_equal(p.download(0), GOLD[f"{name}/blur"])
    assert bits_equal(p.download(1), GOLD[f"{name}/down1"])       # pyrDown of the RAW image
    assert bits_equal(p.download(2), GOLD[f"{name}/down2"])
    d = ctx.depth_pyramid(img, 2, smooth=True)
    assert bits_equal(d.download(0), GOLD[f"{name}/median"])


@pytest.mark.parametrize("shape,levels,smooth", [((376, 1241), 4, True), ((376, 1241), 4, False),
                                                 ((480, 640), 4, True), ((77, 333), 3, True),
                                                 ((2160, 3840), 6, True), ((9, 17), 2, True)])
@pytest.mark.parametrize("kind", ["u8", "float"])
def test_image_pyramid_bit_exact_vs_oracle(ctx, shape, levels, smooth, kind):
    rng = np.random.default_rng(shape[0] * 7 + levels)
    img = rng.integers(0, 256, shape).astype(np.float32) if kind == "u8" else (rng.random(shape) * 255).astype(np.float32)
    want = oracle.image_pyramid(img, levels, smooth)
    p = ctx.image_pyramid(img, levels, smooth)
    assert p.levels == levels
    for l in range(levels):
        assert p.level_size(l) == (want[l].shape[1], want[l].shape[0])
        assert bits_equal(p.download(l), want[l]), f"level {l}"


@pytest.mark.parametrize("shape,levels", [((376, 1241), 4), ((2160, 3840), 6), ((77, 333), 3)])
@pytest.mark.parametrize("smooth", [False, True])
def test_depth_pyramid_bit_exact_vs_oracle(ctx, shape, levels, smooth):
    rng = np.random.default_rng(11)
    dep = (rng.random(shape) * (rng.random(shape) > 0.9)).astype(np.float32)
    want = oracle.depth_pyramid(dep, levels, smooth)
    p = ctx.depth_pyramid(dep, levels, smooth)
    for l in range(levels):
        assert bits_equal(p.download(l), want[l]), f"level {l}"


def test_bad_arguments_return_status(ctx):
    from odometry_b200 import capi
    img = np.zeros((8, 8), np.float32)
    with pytest.raises(capi.OdoError):
        ctx.image_pyramid(img, 9)         # more than ODO_MAX_LEVELS
    with pytest.raises(capi.OdoError):
        ctx.image_pyramid(img, 5)         # level smaller than one pixel
