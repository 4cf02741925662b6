"""K3/K4/K5 parity: selection mask and integer disparities bit-exact, refinement vs oracle. Needs a GPU."""
import numpy as np
import pytest

import oracle
from odometry_b200 import capi

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.fixture(scope="module")
def pairs(synth_data):
    out = []
    for sid in (0, 3):
        d = synth_data.pair(sid)
        out.append((d["L1"], d["R1"]))
    rng = np.random.default_rng(2)
    # a hard case: pure noise right image (many failed matches, ties impossible but thresholds bite)
    out.append((out[0][0], rng.integers(0, 256, out[0][0].shape).astype(np.float32)))
    # a flat case: constant images => zero gradients everywhere, nothing selected
    out.append((np.full((376, 1241), 90, np.float32), np.full((376, 1241), 90, np.float32)))
    return out


@pytest.mark.parametrize("i", range(4))
def test_selection_mask_bit_exact(ctx, pairs, i):
    L, R = pairs[i]
    dp, op = capi.default_depth_params(), oracle.depth_params()
    rc, bl, br, val = ctx.select_pixels(L, R, dp)
    orc, obl, obr, oval = oracle.select_pixels(L, R, op)
    assert rc == 0 and orc == 0
    assert bits_equal(bl, obl) and bits_equal(br, obr)
    assert np.array_equal(val, oval)
    assert val.reshape(-1).sum() <= 512 * 80


@pytest.mark.parametrize("i", range(3))
def test_disparity_indices_bit_exact(ctx, pairs, i):
    L, R = pairs[i]
    dp, op = capi.default_depth_params(), oracle.depth_params()
    _, obl, obr, oval = oracle.select_pixels(L, R, op)
    orc, odisp, odep, omatch, ossd = oracle.disparity_search(obl, obr, oval, op)
    rc, disp, dep, match, ssd = ctx.disparity_search(obl, obr, oval, dp)
    assert rc == 0 and orc == 0
    assert np.array_equal(match, omatch)            # integer argmin, first-minimum tie rule
    assert bits_equal(disp, odisp) and bits_equal(dep, odep)
    sel = oval == 1
    assert bits_equal(ssd[sel], ossd[sel])          # the 8-tap SSD in the AVX tree order


@pytest.mark.parametrize("i", range(3))
def test_disparity_integer_path_bit_exact(ctx, pairs, i):
    """Without the cost plane the library may take its exact integer path (u8-sourced frames blur to
    multiples of 1/16): argmin columns, disparities and inverse depths must not change by a bit."""
    L, R = pairs[i]
    dp, op = capi.default_depth_params(), oracle.depth_params()
    _, obl, obr, oval = oracle.select_pixels(L, R, op)
    assert np.array_equal(obl * 16, np.rint(obl * 16))          # the premise of the integer path
    orc, odisp, odep, omatch, ossd = oracle.disparity_search(obl, obr, oval, op)
    rc, disp, dep, match, _ = ctx.disparity_search(obl, obr, oval, dp, want_ssd=False)
    assert rc == 0 and orc == 0
    assert np.array_equal(match, omatch)
    assert bits_equal(disp, odisp) and bits_equal(dep, odep)


def test_disparity_integer_path_extremes(ctx):
    """Saturated contrast: 256*SSD reaches 8*4080^2 (no int32 overflow), costs above 2^16 are where
    the fp32 reference rounds -- all of them rejected by ssd_th on both sides; ties between equal
    integer costs resolve to the first column."""
    h, w = 376, 1241
    rng = np.random.default_rng(5)
    R = (rng.integers(0, 2, (h, w)) * 255).astype(np.float32)
    L = np.roll(R, 7, axis=1).copy()
    L[:, ::3] = 255 - L[:, ::3]
    op, dp = oracle.depth_params(), capi.default_depth_params()
    bl, br = oracle.gaussian_blur3(L), oracle.gaussian_blur3(R)
    val = np.zeros((h, w), np.uint8)
    val[8:368:5, 12:1230:9] = 1
    for th in (900.0, 20000.0, 65535.0):
        op.ssd_th = dp.ssd_th = th
        orc, odisp, odep, omatch, _ = oracle.disparity_search(bl, br, val, op)
        rc, disp, dep, match, _ = ctx.disparity_search(bl, br, val, dp, want_ssd=False)
        assert np.array_equal(match, omatch), th
        assert bits_equal(disp, odisp) and bits_equal(dep, odep)


def test_disparity_falls_back_to_fp32_for_inexact_input(ctx, pairs):
    """Arbitrary float images (not multiples of 1/16): the integer path must switch itself off."""
    L, R = pairs[0]
    rng = np.random.default_rng(9)
    Lf = (L + rng.random(L.shape).astype(np.float32) * 0.37).astype(np.float32)
    Rf = (R + rng.random(R.shape).astype(np.float32) * 0.37).astype(np.float32)
    dp, op = capi.default_depth_params(), oracle.depth_params()
    _, obl, obr, oval = oracle.select_pixels(Lf, Rf, op)
    orc, odisp, odep, omatch, _ = oracle.disparity_search(obl, obr, oval, op)
    rc, disp, dep, match, _ = ctx.disparity_search(obl, obr, oval, dp, want_ssd=False)
    assert np.array_equal(match, omatch) and bits_equal(disp, odisp) and bits_equal(dep, odep)
    # one single inexact pixel is enough
    bl1 = oracle.gaussian_blur3(L)
    br1 = oracle.gaussian_blur3(R)
    br1[200, 600] += np.float32(1.0 / 64)
    _, _, _, oval1 = oracle.select_pixels(L, R, op)
    orc, odisp, odep, omatch, _ = oracle.disparity_search(bl1, br1, oval1, op)
    rc, disp, dep, match, _ = ctx.disparity_search(bl1, br1, oval1, dp, want_ssd=False)
    assert np.array_equal(match, omatch) and bits_equal(disp, odisp) and bits_equal(dep, odep)


def test_disparity_tie_break_first_minimum(ctx):
    """Periodic right image => many exactly equal SSD minima; the smallest column must win."""
    h, w = 376, 1241
    x = np.arange(w)
    row = ((x % 16) * 13 % 97).astype(np.float32)
    R = np.tile(row, (h, 1))
    L = np.roll(R, 5, axis=1)
    dp, op = capi.default_depth_params(ssd_th=1e9), oracle.depth_params(ssd_th=1e9)
    val = np.zeros((h, w), np.uint8)
    val[10:360:7, 40:1200:11] = 1
    orc, odisp, odep, omatch, ossd = oracle.disparity_search(L, R, val, op)
    rc, disp, dep, match, ssd = ctx.disparity_search(L, R, val, dp)
    assert np.array_equal(match, omatch) and bits_equal(disp, odisp)
    assert (omatch[val == 1] < 40).mean() > 0.9     # the period makes early columns win
    # same through the integer path (ssd_th must stay below 2^16 for it to be eligible)
    dp, op = capi.default_depth_params(ssd_th=60000.0), oracle.depth_params(ssd_th=60000.0)
    orc, odisp, odep, omatch, ossd = oracle.disparity_search(L, R, val, op)
    rc, disp, dep, match, _ = ctx.disparity_search(L, R, val, dp, want_ssd=False)
    assert np.array_equal(match, omatch) and bits_equal(disp, odisp)


@pytest.mark.parametrize("i", range(2))
def test_refinement_matches_oracle(ctx, pairs, i):
    L, R = pairs[i]
    dp, op = capi.default_depth_params(), oracle.depth_params()
    _, obl, obr, oval = oracle.select_pixels(L, R, op)
    _, odisp, odep, _, _ = oracle.disparity_search(obl, obr, oval, op)
    orc, rdep, rval, ost = oracle.depth_optimize(L, R, odep, oval, op)
    rc, gdep, gval, gst = ctx.depth_optimize(L, R, odep, oval, dp)
    assert rc == orc == 0
    assert gst.iters == ost.iters
    assert np.array_equal(gval, rval)
    assert gst.n_valid == ost.n_valid == int(rval.sum())
    assert np.allclose(gdep, rdep, rtol=1e-6, atol=1e-9)
    assert gst.cost == pytest.approx(ost.cost, rel=1e-6)


@pytest.mark.parametrize("i", range(4))
def test_compute_depth_end_to_end(ctx, pairs, i):
    L, R = pairs[i]
    dp, op = capi.default_depth_params(), oracle.depth_params()
    orc, oval, odisp, odep, ost = oracle.compute_depth(L, R, op)
    rc, val, disp, dep, st = ctx.compute_depth(L, R, dp)
    assert rc == orc
    assert st.n_selected == ost.n_selected
    assert bits_equal(disp, odisp)
    assert np.array_equal(val, oval)
    assert np.allclose(dep, odep, rtol=1e-6, atol=1e-9)
    assert (st.iters, st.n_valid) == (ost.iters, ost.n_valid)


def test_compute_depth_rejects_tiny_images(ctx):
    dp = capi.default_depth_params()
    rc, *_ = ctx.compute_depth(np.zeros((20, 30), np.float32), np.zeros((20, 30), np.float32), dp)
    assert rc == -1


@pytest.mark.parametrize("case", ["scene", "repetitive", "flat_rows", "short_rows"])
def test_pruned_scan_equals_plain_scan(ctx, pairs, case, monkeypatch):
    """The integer path scans downwards and rejects chunks from a byte-domain lower bound; the plain
    ascending scan stays available behind ODO_DISP_NOPRUNE.  Both must give the oracle's columns bit for bit, also where
    pruning has nothing to hold on to: periodic texture (many equal costs: the FIRST minimum must win), rows whose
    candidates all cost the same, and pixels so close to the left border that no full chunk exists."""
    h, w = 376, 1241
    op, dp = oracle.depth_params(), capi.default_depth_params()
    rng = np.random.default_rng(11)
    if case == "scene":
        L, R = pairs[1]
        _, bl, br, val = oracle.select_pixels(L, R, op)
    else:
        if case == "repetitive":
            period = rng.integers(0, 256, 16).astype(np.float32)
            R = np.tile(period, (h, w // 16 + 1))[:, :w].copy()
            R += rng.integers(0, 2, (h, 1)).astype(np.float32)      # rows differ, columns repeat every 16
            L = np.roll(R, 5, axis=1)
        elif case == "flat_rows":
            R = np.repeat(rng.integers(0, 256, (h, 1)).astype(np.float32), w, axis=1)
            L = R.copy()
        else:
            R = rng.integers(0, 256, (h, w)).astype(np.float32)
            L = np.roll(R, 3, axis=1)
        bl, br = oracle.gaussian_blur3(L), oracle.gaussian_blur3(R)
        val = np.zeros((h, w), np.uint8)
        if case == "short_rows":
            val[4:h - 4, 4:40:3] = 1                                 # candidates [4, x): fewer than one chunk
            val[10, 4] = 1                                           # no candidate at all
        else:
            val[4:h - 4:3, 6:w - 4:37] = 1
    orc, odisp, odep, omatch, _ = oracle.disparity_search(bl, br, val, op)
    out = {}
    for mode in ("prune", "plain"):
        if mode == "plain":
            monkeypatch.setenv("ODO_DISP_NOPRUNE", "1")
        else:
            monkeypatch.delenv("ODO_DISP_NOPRUNE", raising=False)
        rc, disp, dep, match, _ = ctx.disparity_search(bl, br, val, dp, want_ssd=False)
        assert rc == 0 and orc == 0
        assert np.array_equal(match, omatch), f"{case}/{mode}"
        assert bits_equal(disp, odisp) and bits_equal(dep, odep), f"{case}/{mode}"
        out[mode] = match
    assert np.array_equal(out["prune"], out["plain"])


def test_refinement_state_in_shared_or_global_memory(ctx, pairs, monkeypatch):
    """k_refine keeps {jtwj, b} of its slots in shared memory and falls back to a global array for very large batches;
    both placements must give the same bits (the fallback is forced here through its test switch)."""
    L, R = pairs[0]
    dp = capi.default_depth_params()
    monkeypatch.delenv("ODO_REFINE_GLOBAL_STATE", raising=False)
    a = ctx.compute_depth(L, R, dp)
    monkeypatch.setenv("ODO_REFINE_GLOBAL_STATE", "1")
    b = ctx.compute_depth(L, R, dp)
    assert a[0] == b[0] == 0
    assert np.array_equal(a[1], b[1]) and bits_equal(a[2], b[2]) and bits_equal(a[3], b[3])
    assert (a[4].iters, a[4].n_valid, a[4].cost) == (b[4].iters, b[4].n_valid, b[4].cost)
