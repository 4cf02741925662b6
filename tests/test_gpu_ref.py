"""Device results against the REFERENCE'S OWN code (oracle/_ref, see tests/test_oracle_ref.py), through the
C-ABI.  Needs a GPU and the prebuilt oracle/_ref/libodoref.so (it travels with the snapshot).

Integer work (selection mask, disparities, residual sets) and the pyramids are bit-exact; the N-term sums are
compared at the north-star 1e-5 relative; the converged pose at the bound the association of those sums
allows (tests/test_oracle_ref.py::test_solve_wide_vs_reference_association).
"""
import math

import numpy as np
import pytest

import oracle
from oracle import ref
from odometry_b200 import capi

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")]
W, H, LEVELS = 1241, 376, 4


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


@pytest.fixture(scope="module")
def pair(synth_data):
    d = synth_data.pair(0)
    rc, val, disp, dep, st = ref.compute_depth(d["L1"], d["R1"], oracle.depth_params())
    assert rc == 0
    return dict(d, val=val, disp=disp, dep=dep, st=st)


@pytest.mark.parametrize("smooth", [True, False])
def test_image_pyramid_vs_reference(ctx, pair, smooth):
    want = oracle.split_levels(ref.image_pyramid(pair["L2"], LEVELS, smooth), W, H, LEVELS)
    got = ctx.image_pyramid(pair["L2"], LEVELS, smooth)
    for l in range(LEVELS):
        assert np.array_equal(bits(got.download(l)), bits(want[l])), f"level {l}"


def test_depth_pyramid_vs_reference(ctx, pair):
    want = oracle.split_levels(ref.depth_pyramid(pair["dep"], LEVELS, False), W, H, LEVELS)
    got = ctx.depth_pyramid(pair["dep"], LEVELS, False)
    for l in range(LEVELS):
        assert np.array_equal(bits(got.download(l)), bits(want[l])), f"level {l}"


def test_compute_depth_vs_reference(ctx, pair):
    rc, val, disp, dep, st = ctx.compute_depth(pair["L1"], pair["R1"], capi.default_depth_params())
    assert rc == 0
    assert np.array_equal(val, pair["val"])                  # validity mask after the refinement's filters
    assert np.array_equal(disp, pair["disp"])                # integer disparities
    assert np.allclose(dep, pair["dep"], rtol=1e-6, atol=0)  # refined inverse depth
    assert st.iters == pair["st"].iters and st.n_valid == pair["st"].n_valid


def test_disparity_stage_vs_reference(ctx, pair):
    p = oracle.depth_params()
    rrc, rval, rdisp, rdep = ref.disparity(pair["L1"], pair["R1"], p)
    rc, bl, br, val = ctx.select_pixels(pair["L1"], pair["R1"], capi.default_depth_params())
    assert rc == 0 and np.array_equal(val, rval)
    rc, disp, dep, match, ssd = ctx.disparity_search(bl, br, val, capi.default_depth_params())
    assert rc == 0 and np.array_equal(disp, rdisp) and np.array_equal(bits(dep), bits(rdep))


@pytest.mark.parametrize("robust", [0, 1, 2])
def test_linearize_vs_reference(ctx, pair, robust):
    p1, d1, p2 = (ctx.image_pyramid(pair["L1"], LEVELS, True), ctx.depth_pyramid(pair["dep"], LEVELS, False),
                  ctx.image_pyramid(pair["L2"], LEVELS, True))
    o1 = oracle.split_levels(ref.image_pyramid(pair["L1"], LEVELS, True), W, H, LEVELS)
    o2 = oracle.split_levels(ref.image_pyramid(pair["L2"], LEVELS, True), W, H, LEVELS)
    od = oracle.split_levels(ref.depth_pyramid(pair["dep"], LEVELS, False), W, H, LEVELS)
    Tgt = oracle.se3_rewrap(pair["T_gt"].astype(np.float32))
    for level in range(LEVELS):
        for T in (np.eye(4, dtype=np.float32), Tgt):
            got = ctx.lm_linearize(p1, d1, p2, level, T, robust=robust, huber_delta=28.0)
            r = ref.lm_linearize(o1[level], o2[level], od[level], level, T, robust, 28.0, oracle.SUM_FLOAT_SEQ)
            assert got["rc"] == 0 and got["N"] == r["N"]
            # the reference's own fp32 sequential sums carry an accumulation error that grows with N
            tol = max(2e-5, 2e-9 * r["N"]) * (3 if robust == 2 else 1)
            assert rel_err(got["H"], r["H"]) < tol
            assert rel_err(got["g"], r["g"]) < tol
            assert abs(got["E"] - r["E"]) < tol * abs(r["E"])


def test_solve_vs_reference(ctx, pair):
    p1, d1, p2 = (ctx.image_pyramid(pair["L1"], LEVELS, True), ctx.depth_pyramid(pair["dep"], LEVELS, False),
                  ctx.image_pyramid(pair["L2"], LEVELS, True))
    rc, T, st = ctx.lm_solve(p1, d1, p2, capi.default_lm_params(robust=1))
    rrc, rT = ref.lm_solve(pair["L1"], pair["dep"], pair["L2"], LEVELS, oracle.lm_params(robust=1, sum_mode=oracle.SUM_FLOAT_SEQ))
    assert rc == rrc == 0
    d = T.astype(np.float64) @ np.linalg.inv(rT.astype(np.float64))
    ang = math.acos(max(-1.0, min(1.0, (np.trace(d[:3, :3]) - 1) / 2)))
    assert ang < 1e-4 and np.linalg.norm(T[:3, 3] - rT[:3, 3]) < 1e-3


def test_tracking_vs_reference_across_a_keyframe_switch(synth_data):
    """The batched device tracker against the loop over the reference's classes, across the first keyframe switch
    (frame 5) and the warm start with the pose to the OLD keyframe that follows it (quirk A.8); the 100-frame version
    is tests/test_gpu_long.py."""
    seq = synth_data.sequence(0, 8)
    c = capi.Context(0)
    trk = capi.Tracker(c, W, H, LEVELS, 1)
    rt = ref.Tracker(W, H, LEVELS, oracle.lm_params(sum_mode=oracle.SUM_FLOAT_SEQ), oracle.depth_params(sum_mode=oracle.SUM_FLOAT_SEQ))
    diffs = []
    for f, (L, R) in enumerate(seq["frames"]):
        Lf, Rf = L.astype(np.float32), R.astype(np.float32)
        rc, pa, pk, kf, st = trk.step([Lf], [Rf])
        rrc, rpa, rpk, rkf, rmag = rt.step(Lf, Rf)
        assert rc == rrc == 0
        assert bool(kf[0]) == rkf
        diffs.append(float(np.abs(pa[0] - rpa).max()))
    # two summation orders of the reference's own sums end a Solve 1-5 mm apart, a few frames further when one of them
    # stops a level earlier (profiles/association_noise_r2.json); the typical frame is within millimetres
    assert np.median(diffs) < 5e-3 and max(diffs) < 0.3, diffs
    trk.close()
    rt.close()
    c.close()
