"""Long-sequence operation (BASELINE configs 3 / 4): >= 100 frames x 4 sequences through the batched device tracker,
across ~20 keyframe switches per sequence (run_odometry_kitti_offline.cpp:198-271, keyframe rule :254-270).

  * GPU vs the oracle in ORC_SUM_WIDE (what the GPU implements): same keyframe decisions on every frame, poses
    within the north-star tolerance (1e-4 rad / 1e-4 m) on every frame;
  * GPU vs the reference's own classes (oracle/_ref, fp32 sequential sums): same keyframe decisions on every
    frame, poses within the drift two legal summation orders accumulate over 20 keyframe hops;
  * against the rendered ground truth: the loop TRACKS (keyframes every 5th frame, final error about 1 % of the
    80 m path -- the reference's own orientation drift from floor() sampling, quirk Q1);
  * batching and sharding leave every sequence's 100 poses bit-identical.
Needs a GPU."""
import math
import threading

import numpy as np
import pytest

import oracle
from oracle import ref
from odometry_b200 import capi, shard

pytestmark = pytest.mark.gpu

W, H, LEVELS = 1241, 376, 4
N_FRAMES, SEQS = 100, (0, 1, 2, 3)


def pose_err(Ta, Tb):
    d = Ta.astype(np.float64) @ np.linalg.inv(Tb.astype(np.float64))
    ang = math.acos(max(-1.0, min(1.0, (np.trace(d[:3, :3]) - 1) / 2)))
    return ang, float(np.linalg.norm(Ta[:3, 3].astype(np.float64) - Tb[:3, 3].astype(np.float64)))


@pytest.fixture(scope="module")
def seqs(synth_data):
    return [synth_data.sequence(s, n_frames=N_FRAMES) for s in SEQS]


def run_gpu(ctx, seqs, ids, n_frames=N_FRAMES):
    """the sequences `ids` as ONE batch through the u8 ingest; returns [frame][(pose_abs, pose_kf, kf)]"""
    t = capi.Tracker(ctx, W, H, LEVELS, len(ids))
    out = []
    for k in range(n_frames):
        rc, pa, pk, kf, st = t.step([seqs[i]["frames"][k][0] for i in ids], [seqs[i]["frames"][k][1] for i in ids])
        assert rc == 0 and (st == 0).all(), f"frame {k}"
        out.append((pa, pk, kf))
    t.close()
    return out


@pytest.fixture(scope="module")
def gpu_batch(ctx, seqs):
    return run_gpu(ctx, seqs, list(range(len(SEQS))))


def run_cpu(make, seqs):
    """one CPU tracker per sequence, one thread each (the ctypes calls release the GIL)"""
    out = [None] * len(seqs)

    def work(i):
        t = make()
        out[i] = [t.step(L.astype(np.float32), R.astype(np.float32)) for L, R in seqs[i]["frames"]]
        t.close()

    ths = [threading.Thread(target=work, args=(i,)) for i in range(len(seqs))]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    return out


def test_long_sequences_match_oracle_wide(gpu_batch, seqs):
    want = run_cpu(lambda: oracle.Tracker(W, H, LEVELS), seqs)
    worst_ang = worst_tr = 0.0
    for i in range(len(SEQS)):
        n_kf = 0
        for k in range(N_FRAMES):
            pa, pk, kf = gpu_batch[k]
            orc, opa, opk, okf, omag = want[i][k]
            assert orc == 0
            assert bool(kf[i]) == okf, f"sequence {i} frame {k}: keyframe decision differs from the oracle"
            n_kf += int(kf[i])
            for got, exp in ((pa[i], opa), (pk[i], opk)):
                ang, tr = pose_err(got, exp)
                worst_ang, worst_tr = max(worst_ang, ang), max(worst_tr, tr)
                assert ang < 1e-4 and tr < 1e-4, f"sequence {i} frame {k}: {ang} rad {tr} m"
        assert 19 <= n_kf <= 21, n_kf         # frame 0 + every 5th frame (4.0 m / 3.3 > 1.1)
    print(f"GPU vs oracle(WIDE) over {len(SEQS)} x {N_FRAMES} frames: worst {worst_ang:.2e} rad, {worst_tr:.2e} m")


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built")
def test_long_sequences_vs_reference_classes(gpu_batch, seqs):
    """Against the reference's own classes in its own arithmetic (fp32 running sums).  The reference's converged pose
    itself moves by 1-5 mm per Solve between legal associations of its float sums (measured:
    profiles/association_noise_r2.json, tools/association_noise.py), the stop rule `err ratio > 0.995` turning a rounding
    difference into one LM step more or less, and the loop's warm starts chain those differences.  What must hold:
    identical keyframe decisions on every frame, the typical frame within the measured single-Solve spread, 95 % of
    the frames within 5 cm (a few LM steps); single frames where one association stops a level early and the other
    does not are further apart (observed: up to 23 cm on the third frame after a keyframe switch) and come back together
    on the next frame."""
    mk = lambda: ref.Tracker(W, H, LEVELS, oracle.lm_params(sum_mode=oracle.SUM_FLOAT_SEQ), oracle.depth_params(sum_mode=oracle.SUM_FLOAT_SEQ))
    want = run_cpu(mk, seqs)
    d_kf, d_abs = [], []
    for i in range(len(SEQS)):
        for k in range(N_FRAMES):
            pa, pk, kf = gpu_batch[k]
            rrc, rpa, rpk, rkf, rmag = want[i][k]
            assert rrc == 0
            assert bool(kf[i]) == rkf, f"sequence {i} frame {k}: keyframe decision differs from the reference"
            ang, tr = pose_err(pk[i], rpk)
            d_kf.append(tr)
            assert ang < 5e-3 and tr < 1.0, f"sequence {i} frame {k} pose_to_keyframe: {ang} rad {tr} m"
            ang, tr = pose_err(pa[i], rpa)
            d_abs.append(tr)
            assert ang < 1e-2 and tr < 1.0, f"sequence {i} frame {k} pose_abs: {ang} rad {tr} m"
    d_kf, d_abs = np.array(d_kf), np.array(d_abs)
    print(f"GPU vs reference classes over {len(SEQS)} x {N_FRAMES} frames: pose_to_keyframe |dt| median {np.median(d_kf):.2e} "
          f"p90 {np.percentile(d_kf, 90):.2e} max {d_kf.max():.2e} m; pose_abs |dt| median {np.median(d_abs):.2e} max {d_abs.max():.2e} m")
    assert np.median(d_kf) < 3e-3          # the measured single-Solve association spread (p90 3.1 mm)
    assert np.percentile(d_kf, 95) < 5e-2
    assert np.median(d_abs) < 2e-2 and d_abs[-1] < 0.2   # ... and the trajectories end together


def test_long_sequences_track_ground_truth(gpu_batch, seqs):
    for i in range(len(SEQS)):
        poses = seqs[i]["poses"]
        P0i = np.linalg.inv(poses[0])
        for k in (4, 24, 49, N_FRAMES - 1):
            gt = (P0i @ poses[k]).astype(np.float32)
            ang, tr = pose_err(gpu_batch[k][0][i], gt)
            dist = float(np.linalg.norm(gt[:3, 3]))
            # the reference's loop drifts quadratically (about 0.7 mrad of orientation per keyframe hop): 1 % at 80 m
            assert tr < max(0.02, 0.02 * dist), f"sequence {i} frame {k}: {tr} m off after {dist} m"
        kfs = [k for k in range(N_FRAMES) if gpu_batch[k][2][i]]
        assert kfs[:4] == [0, 5, 10, 15]


def test_batching_and_sharding_bit_identical_over_long_sequences(ctx, gpu_batch, seqs):
    """every sequence's poses over all 100 frames are the same bits alone, in a shard of two, or in the batch of four"""
    n = len(SEQS)
    for rank in range(2):
        ids = shard.shard_sequences(n, 2, rank)
        part = run_gpu(ctx, seqs, ids)
        for k in range(N_FRAMES):
            for j, s in enumerate(ids):
                assert np.array_equal(part[k][0][j], gpu_batch[k][0][s]), (k, s)
                assert np.array_equal(part[k][1][j], gpu_batch[k][1][s]), (k, s)
                assert part[k][2][j] == gpu_batch[k][2][s]
    alone = run_gpu(ctx, seqs, [3])
    for k in range(N_FRAMES):
        assert np.array_equal(alone[k][0][0], gpu_batch[k][0][3]) and np.array_equal(alone[k][1][0], gpu_batch[k][1][3])
