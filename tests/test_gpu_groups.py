"""The two persistent kernels (k_refine, k_lm_solve) split a sequence over G co-resident CTAs that meet at a group
barrier (odometry_b200/csrc/group_sync.cuh).  These tests force G = 1 and G > 1 for both and trip the time-bounded
barrier wait (a CTA that never arrives): the call must come back with -1 / status -2, other state must stay usable.
Needs a GPU."""
import os
from contextlib import contextmanager

import numpy as np
import pytest

from odometry_b200 import capi

pytestmark = pytest.mark.gpu

W, H, LEVELS = 1241, 376, 4


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module")
def pair(synth_data):
    return synth_data.pair(0)


def test_refine_group_sizes_agree(ctx, pair):
    p = capi.default_depth_params()
    base = ctx.compute_depth(pair["L1"], pair["R1"], p)            # batch 1: G = cap / 256 CTAs (> 1)
    assert base[0] == 0
    for g in (1, 2, 7):
        with env(ODO_REFINE_CTAS_PER_SEQ=g):
            rc, val, disp, dep, st = ctx.compute_depth(pair["L1"], pair["R1"], p)
        assert rc == 0 and st.iters == base[4].iters and st.n_valid == base[4].n_valid
        assert np.array_equal(val, base[1]) and np.array_equal(disp, base[2])
        assert np.allclose(dep, base[3], rtol=1e-6, atol=0)      # only the fp64 association of the error sum differs


def test_lm_group_sizes_agree(ctx, pair):
    rc, val, disp, dep, st = ctx.compute_depth(pair["L1"], pair["R1"], capi.default_depth_params())
    p1, d1, p2 = ctx.image_pyramid(pair["L1"], LEVELS, True), ctx.depth_pyramid(dep, LEVELS, False), ctx.image_pyramid(pair["L2"], LEVELS, True)
    base = ctx.lm_solve(p1, d1, p2, capi.default_lm_params())      # G = 16 for one KITTI frame: one cluster
    lin = ctx.lm_linearize(p1, d1, p2, 0, np.eye(4, dtype=np.float32))
    assert base[0] == 0
    # G = 1; G = 3 as one thread-block cluster (DSMEM reduce); G = 3 and 29 on the global-memory barrier (cooperative launch)
    for g, no_cluster in ((1, 0), (3, 0), (3, 1), (29, 1)):
        kw = dict(ODO_LM_CTAS_PER_SEQ=g, **({"ODO_GROUP_NO_CLUSTER": 1} if no_cluster else {}))
        with env(**kw):
            rc, T, s2 = ctx.lm_solve(p1, d1, p2, capi.default_lm_params())
            l2 = ctx.lm_linearize(p1, d1, p2, 0, np.eye(4, dtype=np.float32))
        assert rc == 0 and list(s2.iters) == list(base[2].iters), (g, no_cluster)
        assert np.abs(T - base[1]).max() < 1e-6
        assert l2["N"] == lin["N"] and np.abs(l2["H"] - lin["H"]).max() <= 1e-12 * np.abs(lin["H"]).max()


def test_barrier_timeout_is_reported_not_hung(ctx, pair):
    """ODO_TEST_LOSE_CTA makes the last CTA of every group leave at once: the others wait ODO_BARRIER_TIMEOUT_MS and give
    up.  The C-ABI returns -1, the statistics carry -2 (not the reference's own failure, -1), Solve returns the failure
    pose, and the next call on the same context works."""
    dp = capi.default_depth_params()
    with env(ODO_TEST_LOSE_CTA=1, ODO_BARRIER_TIMEOUT_MS=20):
        rc, val, disp, dep, st = ctx.compute_depth(pair["L1"], pair["R1"], dp)
    assert rc == -1 and st.status == -2
    rc, val, disp, dep, st = ctx.compute_depth(pair["L1"], pair["R1"], dp)
    assert rc == 0 and st.status == 0
    p1, d1, p2 = ctx.image_pyramid(pair["L1"], LEVELS, True), ctx.depth_pyramid(dep, LEVELS, False), ctx.image_pyramid(pair["L2"], LEVELS, True)
    with env(ODO_TEST_LOSE_CTA=1, ODO_BARRIER_TIMEOUT_MS=20):
        rc, T, s = ctx.lm_solve(p1, d1, p2, capi.default_lm_params())
    want = np.eye(4, dtype=np.float32)
    want[:, 3] = 0
    assert rc == -1 and s.status == -2 and np.array_equal(T, want)
    rc, T, s = ctx.lm_solve(p1, d1, p2, capi.default_lm_params())
    assert rc == 0 and s.status == 0


def test_tracker_survives_a_barrier_timeout(ctx, synth_data):
    """a timed-out frame reports status -2 for its sequences and does NOT end them (the reference's `break` is reserved
    for its own depth failure): the following frames track on, against the old keyframe"""
    seq = synth_data.sequence(0, n_frames=8)
    fr = seq["frames"]
    t = capi.Tracker(ctx, W, H, LEVELS, 2)          # 2 sequences: G > 1 for both persistent kernels
    ref = capi.Tracker(ctx, W, H, LEVELS, 2)
    for k in range(3):
        a = t.step([fr[k][0], fr[k][0]], [fr[k][1], fr[k][1]])
        b = ref.step([fr[k][0], fr[k][0]], [fr[k][1], fr[k][1]])
        assert a[0] == b[0] == 0
    with env(ODO_TEST_LOSE_CTA=1, ODO_BARRIER_TIMEOUT_MS=20):
        rc, pa, pk, kf, st = t.step([fr[3][0], fr[3][0]], [fr[3][1], fr[3][1]])
    assert rc == -1 and (st == -2).all() and not kf.any()
    rc, pa, pk, kf, st = t.step([fr[4][0], fr[4][0]], [fr[4][1], fr[4][1]])
    assert rc == 0 and (st == 0).all()
    ref.step([fr[3][0], fr[3][0]], [fr[3][1], fr[3][1]])
    b = ref.step([fr[4][0], fr[4][0]], [fr[4][1], fr[4][1]])
    # same keyframe (frame 0) in both; the warm start differs (the lost frame left none: 1.6 m to go instead of 0.8 m),
    # so the two solves stop a few LM steps apart
    assert np.abs(pa - b[1]).max() < 0.15
    t.close()
    ref.close()
