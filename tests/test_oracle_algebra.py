"""Oracle SE(3) / 6x6 solver pieces against independent numpy / scipy evaluations -- CPU only."""
import numpy as np
import pytest
from scipy.linalg import expm

import oracle


def hat6(a):
    v, w = a[:3], a[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = v
    return M


@pytest.mark.parametrize("seed", range(5))
def test_se3_exp_matches_matrix_exponential(seed):
    rng = np.random.default_rng(seed)
    a = rng.normal(size=6) * np.array([0.3, 0.3, 0.3, 0.05, 0.05, 0.05])
    T = oracle.se3_exp(a)
    assert np.allclose(T, expm(hat6(a)), atol=2e-6)
    assert np.array_equal(T[3], [0, 0, 0, 1])


def test_se3_exp_small_angle_branch():
    a = np.array([0.1, -0.2, 0.3, 1e-7, -2e-7, 1e-7])
    assert np.allclose(oracle.se3_exp(a), expm(hat6(a)), atol=1e-6)
    assert np.allclose(oracle.se3_exp(np.zeros(6)), np.eye(4))


def test_rewrap_is_identity_on_rotations():
    a = np.array([0.5, 0.1, -0.3, 0.02, -0.04, 0.01])
    T = oracle.se3_exp(a)
    assert np.allclose(oracle.se3_rewrap(T), T, atol=3e-7)
    # trace <= 0 branch of Eigen's Quaternion(Matrix3)
    T2 = oracle.se3_exp(np.array([0, 0, 0, 3.0, 0.2, 0.1]))
    assert np.trace(T2[:3, :3]) < 0
    assert np.allclose(oracle.se3_rewrap(T2), T2, atol=1e-6)


def test_mat4_mul():
    rng = np.random.default_rng(1)
    A, B = rng.normal(size=(4, 4)).astype(np.float32), rng.normal(size=(4, 4)).astype(np.float32)
    assert np.allclose(oracle.mat4_mul(A, B), A.astype(np.float64) @ B.astype(np.float64), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("seed", range(6))
def test_solve6_matches_numpy(seed):
    rng = np.random.default_rng(seed)
    J = rng.normal(size=(40, 6)) * np.array([1, 1, 1, 30, 30, 30])   # badly scaled like the LM system
    A = (J.T @ J).astype(np.float32)
    A = A + 0.01 * np.diag(np.diag(A))
    b = rng.normal(size=6).astype(np.float32)
    x = oracle.solve6(A, b)
    ref = np.linalg.solve(A.astype(np.float64), b.astype(np.float64))
    assert np.allclose(x, ref, rtol=2e-3, atol=1e-6)


def test_solve6_rank_deficient_gives_finite_minimum_norm_like_solution():
    A = np.zeros((6, 6), np.float32)
    A[:3, :3] = np.diag([4.0, 2.0, 1.0])
    b = np.array([4, 2, 1, 0, 0, 0], np.float32)
    x = oracle.solve6(A, b)
    assert np.allclose(x, [1, 1, 1, 0, 0, 0], atol=1e-6)


def test_motion_magnitude():
    # run_odometry_kitti_offline.cpp:145: weights (0.1,1,0.1,1,0.1,1)/3.3 on (|ax|,|ay|,|az|,|tx|,|ty|,|tz|)
    T = oracle.se3_exp(np.array([0, 0, 0, 0, 0.3, 0], np.float32))
    T[:3, 3] = [0.5, -0.25, 4.0]
    want = (0.3 * 1.0 + 0.5 * 1.0 + 0.25 * 0.1 + 4.0 * 1.0) / 3.3
    assert oracle.motion_magnitude(T) == pytest.approx(want, rel=1e-5)
