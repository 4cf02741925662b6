"""Drop-in C++ classes (odometry_b200/dropin): the reference's class interfaces over the C-ABI.

CPU part: the library builds against the reference's class declarations and exports every public
member the reference's callers use; without a device it fails loudly.  GPU part: a C++ program that
uses the classes the way run_odometry_kitti_offline.cpp does is compared with the CPU oracle."""
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, LEVELS = 1241, 376, 4


@pytest.fixture(scope="module")
def dropin():
    from odometry_b200.dropin import build
    return build.build()


def test_dropin_exports_the_reference_class_interface(dropin):
    lib, _ = dropin
    syms = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    wanted = [
        "odometry::ImagePyramid::ImagePyramid(int, cv::Mat const&, bool)",          # include/image_pyramid.h:24
        "odometry::ImagePyramid::GetPyramidImage(int) const",                        # :36
        "odometry::DepthPyramid::DepthPyramid(int, cv::Mat const&, bool)",          # :51
        "odometry::DepthPyramid::GetPyramidDepth(int) const",                        # :63
        "odometry::DepthEstimator::ComputeDepth(cv::Mat const&, cv::Mat const&, cv::Mat&, cv::Mat&, cv::Mat&)",
        "odometry::DepthEstimator::ReportStatus()",                                  # include/depth_estimate.h:51,54
        "odometry::LevenbergMarquardtOptimizer::Solve(odometry::ImagePyramid const&, odometry::DepthPyramid const&, "
        "odometry::ImagePyramid const&)",                                            # include/lm_optimizer.h:44
        "odometry::LevenbergMarquardtOptimizer::ShowReport()",
        "odometry::LevenbergMarquardtOptimizer::Reset(Eigen::Matrix<float, 4, 4, 0> const&, float)",
        "odometry::GaussianImagePyramidNaive(int, cv::Mat const&, std::vector<cv::Mat",
        "odometry::MedianDepthPyramidNaive(int, cv::Mat const&, std::vector<cv::Mat",
    ]
    for w in wanted:
        assert w in syms, f"missing from libodometry_dropin.so: {w}"
    # constructors with the reference's argument lists
    assert "odometry::DepthEstimator::DepthEstimator(float, float, float, float, float, float, float, float, int, int, " in syms
    assert "odometry::LevenbergMarquardtOptimizer::LevenbergMarquardtOptimizer(float, float, std::vector<int" in syms


def test_dropin_has_no_cpu_fallback(dropin, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _, driver = dropin
    np.zeros((1, 2, H, W), np.uint8).tofile(tmp_path / "f.u8")
    r = subprocess.run([driver, "units", str(tmp_path / "f.u8"), str(W), str(H), str(tmp_path / "o.bin")],
                       capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stdout


def _level_sizes():
    return [(W >> l, H >> l) for l in range(LEVELS)]


@pytest.mark.gpu
def test_dropin_units_match_oracle(dropin, synth_data, tmp_path):
    """ComputeDepth, both pyramid classes (incl. a copy), the free builders and the inlines."""
    _, driver = dropin
    L, R = synth_data.sequence(0, n_frames=2)["frames"][0]
    np.stack([L, R]).astype(np.uint8).tofile(tmp_path / "f.u8")
    r = subprocess.run([driver, "units", str(tmp_path / "f.u8"), str(W), str(H), str(tmp_path / "o.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "o.bin", "rb").read()
    npx = W * H
    off = 0

    def take(n, dt):
        nonlocal off
        a = np.frombuffer(raw, dt, n, off)
        off += n * np.dtype(dt).itemsize
        return a

    val = take(npx, np.uint8).reshape(H, W)
    disp = take(npx, np.float32).reshape(H, W)
    dep = take(npx, np.float32).reshape(H, W)
    Lf, Rf = L.astype(np.float32), R.astype(np.float32)
    orc, oval, odisp, odep, _ = oracle.compute_depth(Lf, Rf, oracle.depth_params())
    assert orc == 0
    assert np.array_equal(val, oval) and np.array_equal(disp, odisp)
    assert np.allclose(dep, odep, rtol=1e-6, atol=1e-9)
    want_img = oracle.image_pyramid(Lf, LEVELS, True)
    want_dep = oracle.depth_pyramid(dep, LEVELS, False)
    for l, (w, h) in enumerate(_level_sizes()):
        got = take(w * h, np.float32).reshape(h, w)
        assert np.array_equal(got.view(np.uint32), want_img[l].view(np.uint32)), f"image level {l}"
    for l, (w, h) in enumerate(_level_sizes()):
        got = take(w * h, np.float32).reshape(h, w)
        assert np.array_equal(got, want_dep[l]), f"depth level {l}"
    want_img = oracle.image_pyramid(Lf, LEVELS, False)
    want_dep = oracle.depth_pyramid(dep, LEVELS, True)
    for l, (w, h) in enumerate(_level_sizes()):
        got = take(w * h, np.float32).reshape(h, w)
        assert np.array_equal(got.view(np.uint32), want_img[l].view(np.uint32)), f"free image level {l}"
    for l, (w, h) in enumerate(_level_sizes()):
        got = take(w * h, np.float32).reshape(h, w)
        assert np.array_equal(got, want_dep[l]), f"free depth level {l}"
    inl = take(8, np.float32)
    assert inl[0] == np.float32(oracle.get_cx_level(607.1928, 3))
    # ReprojectToCameraFrame / WarpPixel at level 2, evaluated here in numpy
    f = float(np.float32(718.856)) / 4.0
    cx, cy = np.float32(oracle.get_cx_level(607.1928, 2)), np.float32(oracle.get_cx_level(185.2157, 2))
    X = np.float32(float(np.float32(7.5) * (np.float32(100.0) - cx)) / f)
    Y = np.float32(float(np.float32(7.5) * (np.float32(40.0) - cy)) / f)
    assert inl[1] == X and inl[2] == Y
    T = np.eye(4, dtype=np.float64)
    T[0, 3], T[2, 3], T[0, 2], T[2, 0] = np.float32(0.05), np.float32(-0.3), np.float32(0.01), np.float32(-0.01)
    q = T @ np.array([X, Y, 7.5, 1.0])
    assert inl[3] == pytest.approx(f * q[0] / q[2] + float(cx), rel=1e-6)
    assert inl[4] == pytest.approx(f * q[1] / q[2] + float(cy), rel=1e-6)
    assert inl[5] == 0.0
    lvl2 = oracle.image_pyramid(Lf, LEVELS, True)[2]
    assert inl[6] == np.float32(0.5) * (lvl2[40, 101] - lvl2[40, 99])
    assert inl[7] == np.float32(0.5) * (lvl2[41, 100] - lvl2[39, 100])


@pytest.mark.gpu
def test_dropin_tracking_loop_matches_oracle(dropin, synth_data, tmp_path):
    """The reference's main loop (keyframes by value in a tuple, Solve / ComputeDepth / Reset per frame)
    written against the drop-in classes reproduces the oracle tracker."""
    _, driver = dropin
    n = 7
    frames = synth_data.sequence(0, n_frames=8)["frames"][:n]
    np.stack([np.stack([L, R]) for L, R in frames]).astype(np.uint8).tofile(tmp_path / "f.u8")
    r = subprocess.run([driver, "track", str(tmp_path / "f.u8"), str(W), str(H), str(n), str(tmp_path / "o.bin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rec = np.dtype([("pose", np.float32, 16), ("pkf", np.float32, 16), ("kf", np.int32), ("status", np.int32),
                    ("motion", np.float32), ("n_valid", np.int32)])
    got = np.fromfile(tmp_path / "o.bin", rec)
    assert len(got) == n
    t = oracle.Tracker(W, H, LEVELS)
    n_kf = 0
    for k, (L, R) in enumerate(frames):
        orc, opa, opk, okf, omag = t.step(L.astype(np.float32), R.astype(np.float32))
        assert orc == 0 and got["status"][k] == 0
        assert bool(got["kf"][k]) == okf, f"frame {k}"
        n_kf += int(okf)
        assert np.allclose(got["pose"][k].reshape(4, 4).T, opa, atol=1e-4), f"frame {k}"
        assert np.allclose(got["pkf"][k].reshape(4, 4).T, opk, atol=1e-4), f"frame {k}"
        if k > 0:
            assert got["motion"][k] == pytest.approx(omag, abs=1e-4)
        assert got["n_valid"][k] >= 500
    assert n_kf >= 2
    t.close()
