"""The rejection rule of the pruned disparity scan (odometry_b200/csrc/depth.cu, k_disparity<true, pitch, true>) restated
in numpy and attacked on the CPU: a candidate may only be rejected if its exact cost is STRICTLY larger than the best cost
known so far -- for every pattern, every candidate, and with every rounding the kernel applies to the threshold.

    l8 = q >> 4,  S = sum (l8 - r8)^2,  256 SSD >= 256 (S - 2 sqrt(8 S))
    reject  <=>  dot - (Er8 >> 1)  <  thr8,   thr8 = (El8 - int(sthr) - 1) >> 1,
    sthr = min((2.8285 + x * rsqrt(x))^2 * 1.0001 + 2, 1e9),  x = 8 + best / 256   (float32)
"""
import numpy as np
import pytest


def kernel_threshold(El8, best_cost):
    """thr8 exactly as refresh() computes it (float32 arithmetic, rsqrt replaced by its worst case: 2 ulp low or high)"""
    x = np.float32(8.0) + np.float32(best_cost) * np.float32(1.0 / 256.0)
    out = []
    for fudge in (1.0 - 3e-7, 1.0, 1.0 + 3e-7):       # MUFU.RSQ is accurate to about 2^-22
        rs = np.float32(np.float32(1.0) / np.sqrt(x, dtype=np.float32) * np.float32(fudge))
        rt = np.float32(2.8285) + np.float32(x * rs)
        sthr = np.minimum(np.float32(np.float32(rt * rt) * np.float32(1.0001)) + np.float32(2.0), np.float32(1.0e9))
        out.append((El8 - int(sthr) - 1) >> 1)
    return max(out)                                    # the LOOSEST threshold any rounding could produce rejects the most


def costs(lq, rq):
    d = lq.astype(np.int64) - rq.astype(np.int64)
    return (d * d).sum(axis=-1)


def rejected(lq, rq, best_cost):
    l8, r8 = lq >> 4, rq >> 4
    El8 = int((l8.astype(np.int64) ** 2).sum())
    Er8 = (r8.astype(np.int64) ** 2).sum(axis=-1)
    dot = (l8.astype(np.int64) * r8.astype(np.int64)).sum(axis=-1)
    return (dot - (Er8 >> 1)) < kernel_threshold(El8, best_cost)


@pytest.mark.parametrize("seed", range(6))
def test_no_candidate_at_or_below_the_best_cost_is_rejected(seed):
    rng = np.random.default_rng(seed)
    for trial in range(60):
        lq = rng.integers(0, 4081, 8)
        kind = trial % 4
        if kind == 0:      # unrelated candidates
            rq = rng.integers(0, 4081, (4000, 8))
        elif kind == 1:    # near matches: the regime where the bound must not bite
            rq = np.clip(lq + rng.integers(-40, 41, (4000, 8)), 0, 4080)
        elif kind == 2:    # worst case of the byte truncation: every tap differs by 15/16 of a byte step
            rq = np.clip((lq >> 4 << 4) + 15 - ((rng.integers(0, 2, (4000, 8))) * 15) + 16 * rng.integers(-3, 4, (4000, 8)), 0, 4080)
        else:              # flat patterns and saturated values
            lq = np.full(8, rng.integers(0, 4081))
            rq = np.clip(lq + rng.integers(-300, 301, (4000, 8)), 0, 4080)
        c = costs(lq, rq)
        for best in (int(c.min()), int(np.percentile(c, 5)), int(np.median(c)), 0, 900 * 256, 8 * 4080 * 4080):
            rej = rejected(lq, rq, best)
            assert not (rej & (c <= best)).any(), (seed, trial, best)


def test_the_bound_actually_prunes():
    """sanity of the other direction: against a good match, unrelated candidates are rejected almost always"""
    rng = np.random.default_rng(99)
    lq = rng.integers(0, 4081, 8)
    rq = rng.integers(0, 4081, (20000, 8))
    assert rejected(lq, rq, 50 * 256).mean() > 0.99
