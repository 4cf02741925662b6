"""K6/K7 parity: LM linearisation and the full coarse-to-fine solve, device vs oracle. Needs a GPU.

Tolerances (BASELINE.json north_star): per-iteration JtJ / Jtr within 1e-5 relative (summation
order differs), converged pose within 1e-4 rad / 1e-4 m.
"""
import math

import numpy as np
import pytest

import oracle
from odometry_b200 import capi

pytestmark = pytest.mark.gpu
REL = 1e-5


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def pose_err(Ta, Tb):
    d = Ta.astype(np.float64) @ np.linalg.inv(Tb.astype(np.float64))
    ang = math.acos(max(-1.0, min(1.0, (np.trace(d[:3, :3]) - 1) / 2)))
    return ang, float(np.linalg.norm(Ta[:3, 3].astype(np.float64) - Tb[:3, 3].astype(np.float64)))


@pytest.fixture(scope="module")
def scene(synth_data):
    d = synth_data.pair(0)
    p = oracle.depth_params()
    rc, val, disp, dep, st = oracle.compute_depth(d["L1"], d["R1"], p)
    assert rc == 0
    inv_exact = (1.0 / d["z1"]).astype(np.float32)
    return dict(d, dep_stereo=dep, dep_mask=np.where(val == 1, inv_exact, 0).astype(np.float32), dep_dense=inv_exact)


def build(ctx, scene, depth_key, smooth=True, levels=4):
    return (ctx.image_pyramid(scene["L1"], levels, smooth), ctx.depth_pyramid(scene[depth_key], levels, False),
            ctx.image_pyramid(scene["L2"], levels, smooth))


def oracle_pyrs(scene, depth_key, smooth=True, levels=4):
    return (oracle.image_pyramid(scene["L1"], levels, smooth), oracle.depth_pyramid(scene[depth_key], levels, False),
            oracle.image_pyramid(scene["L2"], levels, smooth))


POSES = {"identity": np.eye(4, dtype=np.float32), "gt": None, "perturbed": "perturbed"}


@pytest.mark.parametrize("depth_key", ["dep_mask", "dep_dense", "dep_stereo"])
@pytest.mark.parametrize("robust", [0, 1, 2])
@pytest.mark.parametrize("pose", ["identity", "gt", "perturbed"])
def test_linearize_matches_oracle_every_level(ctx, scene, depth_key, robust, pose):
    if pose == "identity":
        T = np.eye(4, dtype=np.float32)
    elif pose == "gt":
        T = oracle.se3_rewrap(scene["T_gt"].astype(np.float32))
    else:
        T = oracle.mat4_mul(oracle.se3_exp(np.array([0.05, -0.02, 0.1, 0.004, -0.006, 0.003], np.float32)),
                            oracle.se3_rewrap(scene["T_gt"].astype(np.float32)))
    p1, d1, p2 = build(ctx, scene, depth_key)
    o1, od, o2 = oracle_pyrs(scene, depth_key)
    for level in range(4):
        got = ctx.lm_linearize(p1, d1, p2, level, T, robust=robust, huber_delta=28.0)
        wide = oracle.lm_linearize(o1[level], o2[level], od[level], level, T, robust, 28.0, oracle.SUM_WIDE)
        seq = oracle.lm_linearize(o1[level], o2[level], od[level], level, T, robust, 28.0, oracle.SUM_FLOAT_SEQ)
        assert got["rc"] == 0 and wide["rc"] == 0
        # residual set is integer work: identical count
        assert got["N"] == wide["N"] == seq["N"], f"level {level}"
        # same per-residual factors, fp64 accumulation on both sides: only the summation order differs
        assert rel_err(got["H"], wide["H"]) < 1e-12
        assert rel_err(got["g"], wide["g"]) < 1e-10
        assert abs(got["E"] - wide["E"]) <= 1e-12 * abs(wide["E"])
        if robust == 2:
            assert got["sigma"] == pytest.approx(wide["sigma"], rel=1e-6)
        # against fp32 SEQUENTIAL sums (the worst association the reference's Eigen GEMM could have):
        # the difference is the fp32 accumulation error of that oracle mode itself, which grows with
        # N (measured 1e-5 at N = 24k, 4e-4 at N = 454k between the two oracle modes); at the
        # reference's semi-dense sizes it is the north-star 1e-5 (order-of-summation) tolerance.
        tol = max(2 * REL, 2e-9 * got["N"])
        assert rel_err(got["H"], seq["H"]) < tol
        assert rel_err(got["g"], seq["g"]) < tol


@pytest.mark.parametrize("depth_key", ["dep_mask", "dep_stereo", "dep_dense"])
@pytest.mark.parametrize("robust", [1, 2])
@pytest.mark.parametrize("smooth", [True, False])
def test_solve_matches_oracle(ctx, scene, depth_key, robust, smooth):
    p1, d1, p2 = build(ctx, scene, depth_key, smooth)
    lp = capi.default_lm_params(robust=robust)
    rc, T, st = ctx.lm_solve(p1, d1, p2, lp)
    olp = oracle.lm_params(robust=robust, sum_mode=oracle.SUM_WIDE)
    o1 = oracle.image_pyramid(scene["L1"], 4, smooth, packed=True)
    o2 = oracle.image_pyramid(scene["L2"], 4, smooth, packed=True)
    od = oracle.depth_pyramid(scene[depth_key], 4, False, packed=True)
    orc, oT, ost = oracle.lm_solve(o1, od, o2, 1241, 376, 4, olp)
    assert rc == 0 and orc == 0
    ang, tr = pose_err(T, oT)
    assert ang < 1e-4 and tr < 1e-4, (ang, tr, list(st.iters)[:4], list(ost.iters)[:4])
    assert list(st.iters)[:4] == list(ost.iters)[:4]
    assert list(st.n_res)[:4] == list(ost.n_res)[:4]
    # and it actually converged towards the rendered motion
    ang_gt, tr_gt = pose_err(T, scene["T_gt"].astype(np.float32))
    assert ang_gt < 1e-2 and tr_gt < 0.1


def test_solve_warm_start_and_determinism(ctx, scene):
    p1, d1, p2 = build(ctx, scene, "dep_mask")
    lp = capi.default_lm_params()
    T0 = oracle.se3_exp(np.array([0.0, 0.0, -0.25, 0, 0.005, 0], np.float32))
    a = ctx.lm_solve(p1, d1, p2, lp, T0)
    b = ctx.lm_solve(p1, d1, p2, lp, T0)
    assert np.array_equal(a[1], b[1])      # run-to-run bit identical (fixed-order reductions)
    o1 = oracle.image_pyramid(scene["L1"], 4, True, packed=True)
    o2 = oracle.image_pyramid(scene["L2"], 4, True, packed=True)
    od = oracle.depth_pyramid(scene["dep_mask"], 4, False, packed=True)
    orc, oT, _ = oracle.lm_solve(o1, od, o2, 1241, 376, 4, oracle.lm_params(), T0)
    ang, tr = pose_err(a[1], oT)
    assert ang < 1e-4 and tr < 1e-4


def test_solve_fails_like_reference_without_depth(ctx, scene):
    """No valid depth anywhere => num_residual == 0 => Solve returns identity with a zero last column."""
    p1 = ctx.image_pyramid(scene["L1"], 4, True)
    p2 = ctx.image_pyramid(scene["L2"], 4, True)
    d0 = ctx.depth_pyramid(np.zeros_like(scene["L1"]), 4, False)
    rc, T, st = ctx.lm_solve(p1, d0, p2, capi.default_lm_params())
    want = np.eye(4, dtype=np.float32)
    want[:, 3] = 0
    assert rc == -1 and np.array_equal(T, want)
    orc, oT, _ = oracle.lm_solve(oracle.image_pyramid(scene["L1"], 4, True, packed=True),
                                 np.zeros(oracle.lib().orc_pyramid_floats(1241, 376, 4), np.float32),
                                 oracle.image_pyramid(scene["L2"], 4, True, packed=True), 1241, 376, 4, oracle.lm_params())
    assert orc == -1 and np.array_equal(oT, want)


def test_mismatched_pyramids_rejected(ctx, scene):
    p1 = ctx.image_pyramid(scene["L1"], 4, True)
    p3 = ctx.image_pyramid(scene["L1"][:200, :600], 4, True)
    d1 = ctx.depth_pyramid(scene["dep_mask"], 4, False)
    rc, T, _ = ctx.lm_solve(p1, d1, p3, capi.default_lm_params())
    assert rc == -1


@pytest.mark.parametrize("dense_path", [False, True])
def test_dense_4k_linearize(ctx, dense_path, monkeypatch):
    """Config 5 shape: 3840x2160, dense depth, level 0 and the coarsest level of a 6-level pyramid; through the point
    list (default) and through the list-free evaluation of dense levels (ODO_LM_DENSE=1)."""
    import torch
    from odometry_b200 import synth
    monkeypatch.setenv("ODO_LM_DENSE", "1" if dense_path else "0")
    sc = synth.make_scene(5, 2)
    s = 3840 / 1241.0
    kw = dict(w=3840, h=2160, device="cuda", fx=synth.FX * s, cx=synth.CX * s, cy=synth.CY * s)
    T1 = synth.pose_from_motion()
    T2 = synth.pose_from_motion(0.01, 0.0, 0.1, 0, math.radians(0.1), 0)
    L1, z1 = synth.render(sc, T1, **kw)
    L2, _ = synth.render(sc, T2, **kw)
    L1, L2 = L1.cpu().numpy().astype(np.float32), L2.cpu().numpy().astype(np.float32)
    dep = (1.0 / z1.cpu().numpy()).astype(np.float32)
    p1, d1, p2 = ctx.image_pyramid(L1, 6, True), ctx.depth_pyramid(dep, 6, False), ctx.image_pyramid(L2, 6, True)
    o1, od, o2 = oracle.image_pyramid(L1, 6, True), oracle.depth_pyramid(dep, 6, False), oracle.image_pyramid(L2, 6, True)
    T = np.eye(4, dtype=np.float32)
    for level in (0, 5):
        got = ctx.lm_linearize(p1, d1, p2, level, T, robust=1)
        wide = oracle.lm_linearize(o1[level], o2[level], od[level], level, T, 1, 28.0, oracle.SUM_WIDE)
        assert got["N"] == wide["N"]
        assert rel_err(got["H"], wide["H"]) < 1e-11
        assert rel_err(got["g"], wide["g"]) < 1e-9


_CASE_4K = {}


def _case_4k():
    """frames, oracle pyramids and the oracle's Solve of the config-5 parity case (computed once: ~25 s of CPU)"""
    if not _CASE_4K:
        from odometry_b200 import synth
        sc = synth.make_scene(1000, 2)
        kw = dict(w=3840, h=2160, device="cuda", fx=synth.FX, cx=synth.CX, cy=synth.CY)   # the reference's hard-coded camera
        T1 = synth.pose_from_motion()
        T2 = synth.pose_from_motion(0.0, 0.0, 0.06, 0.0, math.radians(0.03), 0.0)
        L1, z1 = synth.render(sc, T1, **kw)
        L2, _ = synth.render(sc, T2, **kw)
        L1, L2 = L1.cpu().numpy().astype(np.float32), L2.cpu().numpy().astype(np.float32)
        dep = (1.0 / z1.cpu().numpy()).astype(np.float32)
        iters = (10, 20, 30, 30, 30, 30)
        orc, oT, ost = oracle.lm_solve(oracle.image_pyramid(L1, 6, True, packed=True), oracle.depth_pyramid(dep, 6, False, packed=True),
                                       oracle.image_pyramid(L2, 6, True, packed=True), 3840, 2160, 6, oracle.lm_params(max_iters=iters))
        _CASE_4K.update(L1=L1, L2=L2, dep=dep, iters=iters, orc=orc, oT=oT, ost=ost, gt=synth.relative_pose(T1, T2).astype(np.float32))
    return _CASE_4K


@pytest.mark.parametrize("dense_path", [False, True])
def test_dense_4k_solve_matches_oracle(ctx, dense_path, monkeypatch):
    """Config 5: one 3840x2160 frame, 6-level pyramids, all-pixel dense alignment -- the whole coarse-to-fine Solve
    (src/lm_optimizer.cpp:73-160, iters {10,20,30,30,30,30}) against the oracle in ORC_SUM_WIDE, on both evaluation paths
    (point list; list-free evaluation of dense levels, ODO_LM_DENSE=1)."""
    monkeypatch.setenv("ODO_LM_DENSE", "1" if dense_path else "0")
    c = _case_4k()
    p1, d1, p2 = ctx.image_pyramid(c["L1"], 6, True), ctx.depth_pyramid(c["dep"], 6, False), ctx.image_pyramid(c["L2"], 6, True)
    rc, T, st = ctx.lm_solve(p1, d1, p2, capi.default_lm_params(max_iters=c["iters"]))
    assert rc == c["orc"] == 0
    assert list(st.iters)[:6] == list(c["ost"].iters)[:6] and list(st.n_res)[:6] == list(c["ost"].n_res)[:6]
    ang, tr = pose_err(T, c["oT"])
    assert ang < 1e-4 and tr < 1e-4, (ang, tr)
    ang, tr = pose_err(T, c["gt"])
    assert tr < 5e-3 and ang < 1e-3, (ang, tr)
    for p in (p1, d1, p2):
        p.release()
