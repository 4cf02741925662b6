"""The short cuts the kernels take around divisions, square roots and floor() are claimed to be EXACT (same bits as the
expression of the reference they replace), not approximations.  The GPU parity tests sample those claims on rendered
frames; here the arithmetic of each short cut is restated in numpy (IEEE float32 / float64, FMAs emulated exactly) and
attacked on the CPU with random and adversarial operands.

  * floor_proj  (odometry_b200/csrc/lm.cu): floor of WarpPixel's projected coordinate
    (include/image_processing_global.h:50-53) through an fp32 fast path that falls back to the fp64 expression
    near integers;
  * div_rcp     (lm.cu): a / b through a shared correctly rounded reciprocal, used by the back substitution of the
    6x6 solve and by the list-free dense evaluation;
  * div_by_f, f_over_z (lm.cu): the fp64 quotients of ReprojectToCameraFrame / the warp Jacobian
    (include/image_processing_global.h:31-39, src/lm_optimizer.cpp:223-233) through one reciprocal;
  * the squared-magnitude threshold of k_select_tile (depth.cu): `sqrt(v) > th` (src/depth_estimate.cpp:330) as
    `v >= tv`;
  * the integer cost of k_disparity (depth.cu) against the fp32 hadd tree of ComputeSsdPattern8Sse
    (src/depth_estimate.cpp:435-453).
"""
from fractions import Fraction

import numpy as np
import pytest

F32 = np.float32


# ------------------------------------------------------------------------------------------------------------------
# helpers: exact FMAs
# ------------------------------------------------------------------------------------------------------------------
def fma32(a, b, c):
    """RN32(a * b + c) for float32 arrays: the product is exact in fp64 (24 + 24 bits), the fp64 sum is made
    round-to-odd with the TwoSum error term, so the final rounding to fp32 is the only one that counts."""
    a, b, c = a.astype(np.float64), b.astype(np.float64), c.astype(np.float64)
    p = a * b
    s = p + c
    t = s - p
    e = (p - (s - t)) + (c - t)
    even = (s.view(np.int64) & 1) == 0
    toward = np.nextafter(s, np.where(e > 0, np.inf, -np.inf))
    return np.where((e != 0) & even, toward, s).astype(F32)


def fma64(a, b, c):
    """RN64(a * b + c) for Python floats through exact rationals (Fraction -> float is correctly rounded)"""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def test_fma32_emulation_is_exact():
    rng = np.random.default_rng(0)
    a = rng.standard_normal(2000).astype(F32) * F32(1e3)
    b = rng.standard_normal(2000).astype(F32)
    c = (-a * b).astype(F32) * F32(1 + 1e-6)         # heavy cancellation: the case a separate multiply gets wrong
    want = np.array([np.float32(_rn32(Fraction(float(x)) * Fraction(float(y)) + Fraction(float(z))))
                     for x, y, z in zip(a, b, c)], F32)
    assert np.array_equal(fma32(a, b, c), want)


def _rn32(q: Fraction) -> float:
    """a rational rounded ONCE to the nearest float32 (ties to even), normal range"""
    if q == 0:
        return 0.0
    s = -1 if q < 0 else 1
    q = abs(q)
    e = q.numerator.bit_length() - q.denominator.bit_length()
    if Fraction(2) ** e > q:
        e -= 1
    scaled = q / Fraction(2) ** (e - 23)             # in [2^23, 2^24)
    n = scaled.numerator // scaled.denominator
    rem = scaled - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1)):
        n += 1
    return s * float(n) * 2.0 ** (e - 23)


# ------------------------------------------------------------------------------------------------------------------
# floor_proj
# ------------------------------------------------------------------------------------------------------------------
def floor_reference(q, qz, f, c):
    """floor((float)(f * q / qz + c)) with the double intermediates of WarpPixel"""
    u = (f * q.astype(np.float64)) / qz.astype(np.float64) + c.astype(np.float64)
    return np.floor(u.astype(F32))


def floor_fast(q, qz, f, c):
    """the fp32 path of floor_proj; returns (floor, taken): where `taken` is False the kernel evaluates the fp64
    expression instead"""
    ff = F32(f)
    rz = F32(1.0) / qz
    m = ff * (q * rz)
    u = m + c
    fl = np.floor(u)
    fr = u - fl
    tol = (np.abs(m) + np.abs(u)) * F32(4.7683716e-07)
    taken = (fr >= tol) & (fr <= F32(1.0) - tol)
    return fl, taken


@pytest.mark.parametrize("level", [0, 1, 2, 3, 5])
def test_floor_fast_path_agrees_with_the_double_expression(level):
    rng = np.random.default_rng(100 + level)
    f = float(F32(718.856)) / 2.0 ** level
    n = 2_000_000
    taken_total = 0
    for attack in ("uniform", "near_integers", "image_border"):
        qz = rng.uniform(0.3, 120.0, n).astype(F32)
        c = np.full(n, F32(607.1928 / 2 ** level) if attack != "image_border" else F32(185.2157 / 2 ** level), F32)
        if attack == "uniform":
            u_target = rng.uniform(-50.0, 4000.0 / 2 ** level, n)
        elif attack == "near_integers":
            # projected coordinates a few ulps either side of an integer: where a wrong floor would come from
            k = rng.integers(0, max(4000 >> level, 8), n).astype(np.float64)
            u_target = k + rng.choice([-1.0, 1.0], n) * np.exp(rng.uniform(np.log(1e-9), np.log(1e-3), n))
        else:
            u_target = rng.choice([0.0, 1.0, (1241 >> level) - 1.0, float(1241 >> level)], n) + rng.normal(0, 2e-5, n)
        q = ((u_target - c) * qz.astype(np.float64) / f).astype(F32)
        fl, taken = floor_fast(q, qz, f, c)
        want = floor_reference(q, qz, f, c)
        assert np.array_equal(fl[taken], want[taken]), attack
        taken_total += int(taken.sum())
        if attack == "uniform":
            assert taken.mean() > 0.99          # the fall-back is the exception, not the rule
    assert taken_total > 0


def test_floor_fast_path_rejects_non_finite():
    q = np.array([np.nan, np.inf, 1.0, 1.0], F32)
    qz = np.array([1.0, 1.0, np.inf, 1e-38], F32)
    with np.errstate(all="ignore"):
        _, taken = floor_fast(q, qz, 718.856, np.full(4, 600.0, F32))
    assert not taken[0] and not taken[1] and not taken[3]


# ------------------------------------------------------------------------------------------------------------------
# div_rcp
# ------------------------------------------------------------------------------------------------------------------
def div_rcp(a, b):
    rb = F32(1.0) / b
    q0 = a * rb
    rem = fma32(-b, q0, a)
    return fma32(rem, rb, q0)


def _random_f32(rng, n, emin, emax, signed=True):
    m = rng.integers(1 << 23, 1 << 24, n).astype(np.float64)
    x = np.ldexp(m, rng.integers(emin, emax, n) - 23).astype(F32)
    return x * rng.choice(np.array([-1, 1], F32), n) if signed else x


@pytest.mark.parametrize("seed", range(3))
def test_div_rcp_is_the_correctly_rounded_quotient(seed):
    rng = np.random.default_rng(seed)
    n = 1_000_000
    a, b = _random_f32(rng, n, -20, 20), _random_f32(rng, n, -20, 20)
    assert np.array_equal(div_rcp(a, b), a / b)
    # divisors whose mantissa is (nearly) all ones: the worst case of a rounded reciprocal
    b = np.ldexp(((1 << 24) - 1 - rng.integers(0, 64, n)).astype(np.float64), -23).astype(F32)
    a = _random_f32(rng, n, -3, 3)
    assert np.array_equal(div_rcp(a, b), a / b)
    # quotients on and next to rounding midpoints
    mq = rng.integers(1 << 23, 1 << 24, n).astype(np.float64) + 0.5
    mb = rng.integers(1 << 23, 1 << 24, n).astype(np.float64)
    b = (mb * 2.0 ** -23).astype(F32)
    a0 = (mq * mb * 2.0 ** -46).astype(F32)
    for a in (a0, np.nextafter(a0, F32(np.inf)), np.nextafter(a0, F32(-np.inf))):
        assert np.array_equal(div_rcp(a, b), a / b)


# ------------------------------------------------------------------------------------------------------------------
# div_by_f, f_over_z
# ------------------------------------------------------------------------------------------------------------------
def div_by_f(x, f, r):
    q0 = x * r
    rem = fma64(-f, q0, x)
    return fma64(rem, r, q0)


@pytest.mark.parametrize("level", range(6))
def test_div_by_f_is_the_fp64_quotient(level):
    """(float)(x / f): x = z * (x_pix - cx) is a float widened to double, f the level's focal length"""
    rng = np.random.default_rng(200 + level)
    f = float(F32(718.856)) / 2.0 ** level
    r = 1.0 / f
    z = (1.0 / rng.uniform(0.01, 3.0, 3000)).astype(F32)
    xc = (rng.integers(0, 3840 >> level, 3000).astype(F32) - F32(1920.5 / 2 ** level)).astype(F32)
    for x in (z * xc).astype(F32):
        x = float(x)
        assert div_by_f(x, f, r) == x / f


def f_over_z(f, z, d):
    zd, r = float(z), float(d)
    r = fma64(r, fma64(-zd, r, 1.0), r)
    r = fma64(r, fma64(-zd, r, 1.0), r)
    q0 = f * r
    rem = fma64(-zd, q0, f)
    return fma64(rem, r, q0)


@pytest.mark.parametrize("level", [0, 2, 5])
def test_f_over_z_is_the_fp64_quotient(level):
    """(float)(f / (double)z) with z = RN32(1 / d): d seeds the reciprocal of z"""
    rng = np.random.default_rng(300 + level)
    f = float(F32(718.856)) / 2.0 ** level
    d = rng.uniform(0.01, 3.0, 3000).astype(F32)
    d[:200] = np.nextafter(F32(1.0) / rng.integers(1, 100, 200).astype(F32), F32(0))
    z = F32(1.0) / d
    for zi, di in zip(z, d):
        assert f_over_z(f, zi, di) == f / float(zi)


# ------------------------------------------------------------------------------------------------------------------
# squared-magnitude threshold of the selection
# ------------------------------------------------------------------------------------------------------------------
def squared_threshold(th):
    """tv of k_select_tile: the midpoint between th and the next float, squared in fp64 (exact: 25 x 25 bits), rounded UP
    to fp32"""
    nx = (th.view(np.uint32) + np.uint32(1)).view(F32)
    mid = 0.5 * (th.astype(np.float64) + nx.astype(np.float64))
    sq = mid * mid
    tv = sq.astype(F32)
    return np.where(tv.astype(np.float64) < sq, np.nextafter(tv, F32(np.inf)), tv)


@pytest.mark.parametrize("seed", range(3))
def test_squared_threshold_selects_exactly_the_pixels_above_the_threshold(seed):
    rng = np.random.default_rng(400 + seed)
    n = 200_000
    # thresholds as the kernel forms them: a median gradient magnitude plus grad_th, and the extremes
    th = np.concatenate([
        (np.sqrt(rng.integers(0, 2 * 255 * 255, n).astype(F32) * F32(0.25)) + F32(8.0)).astype(F32),
        _random_f32(rng, n, -30, 30, signed=False),
        np.array([0.0, 1.4e-45, 1.1754944e-38, 8.0, 368.624, 1e18], F32)])
    tv = squared_threshold(th)
    with np.errstate(over="ignore"):
        for steps in range(-3, 4):
            v = tv
            for _ in range(abs(steps)):
                v = np.nextafter(v, F32(np.inf) if steps > 0 else F32(0))
            assert np.array_equal(np.sqrt(v) > th, v >= tv), steps
    # magnitudes as the kernel forms them (quarter-integer squares of u8-sourced gradients) around one threshold each
    g = rng.integers(0, 511, (n, 2)).astype(F32) * F32(0.5) * F32(0.5)
    v = g[:, 0] * g[:, 0] + g[:, 1] * g[:, 1]
    t = th[:n]
    assert np.array_equal(np.sqrt(v) > t, v >= squared_threshold(t))


# ------------------------------------------------------------------------------------------------------------------
# integer SSD of the line search
# ------------------------------------------------------------------------------------------------------------------
def ssd_tree_f32(l, r):
    """ComputeSsdPattern8Sse (src/depth_estimate.cpp:435-453): fp32 differences, squares and the hadd tree"""
    s = (l - r) * (l - r)
    return ((s[..., 0] + s[..., 1]) + (s[..., 2] + s[..., 3])) + ((s[..., 4] + s[..., 5]) + (s[..., 6] + s[..., 7]))


@pytest.mark.parametrize("spread", [40, 600, 4080])
def test_integer_ssd_equals_the_fp32_tree_wherever_a_match_is_possible(spread):
    """k_disparity's integer path: blurred u8-sourced pixels are multiples of 1/16, q = 16 * blur, and 256 * SSD is the
    integer sum of (ql - qr)^2.  Below 2^24 (SSD < 65536, far above any usable ssd_th) every partial sum of the fp32 tree
    is exact, so the float cost IS the integer cost / 256 and `first minimum` / `<= ssd_th` decide identically; above it
    the float cost stays above 65535, i.e. such a candidate can neither win against an exact one nor pass the threshold."""
    rng = np.random.default_rng(spread)
    n = 1_000_000
    ql = rng.integers(0, 4081, (n, 8))
    qr = np.clip(ql + rng.integers(-spread, spread + 1, (n, 8)), 0, 4080)
    cost = ((ql - qr) ** 2).sum(axis=1)
    f = ssd_tree_f32(ql.astype(F32) / F32(16), qr.astype(F32) / F32(16))
    small = cost < (1 << 24)
    assert np.array_equal(f[small].astype(np.float64) * 256.0, cost[small].astype(np.float64))
    assert (f[~small] > 65535.0).all()
    if spread == 4080:
        assert (~small).any() and small.any()
