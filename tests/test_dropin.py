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


def test_dropin_exports_the_camera_interface(dropin):
    lib, _ = dropin
    syms = subprocess.run(["nm", "-DC", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    for w in ["odometry::CameraPyramid::CameraPyramid(int, double, double, double, double, double, double, double, double, "
              "double, double, double, int, int)",                                               # include/camera.h:37-38
              "odometry::CameraPyramid::ConfigureCamera(cv::Mat const&, cv::Mat const&, cv::Size const&, int, bool)",  # :56-57
              "odometry::CameraPyramid::UndistortRectify(cv::Mat const&, cv::Mat&, int, int, cv::Scalar const&)",   # :68
              "odometry::SetUpStereoCameraSystem(", "odometry::ReadStereoCalibrationFile("]:   # :126-133
        assert w in syms, f"missing from libodometry_dropin.so: {w}"


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


def _write_camchain(path):
    """the reference's calibration_file/camchain.yaml, re-emitted from the fixture constants"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import camera_patterns as cp
    rows = "\n".join("  - " + repr(list(r)) for r in cp.T_CN_CNM1 + [[0.0, 0.0, 0.0, 1.0]])
    txt = f"""cam0:
  cam_overlaps: [1]
  camera_model: pinhole
  distortion_coeffs: {cp.CAM0_DIST!r}
  distortion_model: radtan
  intrinsics: {cp.CAM0_INTR!r}
  resolution: [640, 482]
  sensor_size: [5.76, 4.29]
  rostopic: /cam0/image_raw
cam1:
  T_cn_cnm1:
{rows}
  cam_overlaps: [0]
  camera_model: pinhole
  distortion_coeffs: {cp.CAM1_DIST!r}
  distortion_model: radtan
  intrinsics: {cp.CAM1_INTR!r}
  resolution: [640, 482]
  sensor_size: [5.76, 4.29]
  rostopic: /cam1/image_raw
"""
    open(path, "w").write(txt)
    return cp


@pytest.mark.gpu
def test_dropin_camera_setup_and_rectification(dropin, tmp_path):
    """test_camera_setup.cpp against the drop-in: calibration file -> CameraPyramid x2 -> ConfigureCamera ->
    UndistortRectify, checked against the cv2 golden vectors (tests/golden/cv2_camera.npz)."""
    _, driver = dropin
    cp = _write_camchain(tmp_path / "camchain.yaml")
    G = np.load(os.path.join(ROOT, "tests", "golden", "cv2_camera.npz"))
    np.concatenate([G["R1"].ravel(), G["P1"].ravel(), G["R2"].ravel(), G["P2"].ravel()]).astype(np.float64).tofile(tmp_path / "rect.f64")
    src8 = cp.raw_pattern_u8()
    np.stack([src8, src8]).tofile(tmp_path / "raw.u8")
    r = subprocess.run([driver, "camera", str(tmp_path / "camchain.yaml"), str(tmp_path / "rect.f64"), str(tmp_path / "raw.u8"),
                        str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "o.bin", "rb").read()
    off = 0

    def take(n, dt):
        nonlocal off
        a = np.frombuffer(raw, dt, n, off)
        off += n * np.dtype(dt).itemsize
        return a

    for k, (K, D) in enumerate((("K0", "D0"), ("K1", "D1"))):
        assert np.array_equal(take(9, np.float64).reshape(3, 3), G[K])
        assert np.array_equal(take(4, np.float64), G[D])
        assert np.array_equal(take(6, np.float64), [5.76, 4.29, 640 / 5.76, 482 / 4.29, 640, 482])
    assert np.array_equal(take(3, np.float64), G["t"].ravel())
    assert np.array_equal(take(9, np.float64).reshape(3, 3), G["R"])
    want = [oracle.camera_intrinsic_pyramid(G[P], 4) for P in ("P1", "P2")]
    for k in range(2):
        assert np.array_equal(take(20, np.float64).reshape(4, 5), want[k])
    for k in range(2):
        assert np.array_equal(take(4, np.float32), want[k][1, :4].astype(np.float32))
    npx = 640 * 480
    left8 = take(npx, np.uint8).reshape(480, 640)
    right8 = take(npx, np.uint8).reshape(480, 640)
    rightf = take(npx, np.float32).reshape(480, 640)
    assert str(cp.digest(left8)) == str(G["left/remap_u8_c0_sha"])
    assert str(cp.digest(right8)) == str(G["right/remap_u8_c0_sha"])
    mx, my = oracle.init_undistort_rectify_map(G["K1"], G["D1"], G["R2"], G["P2"], cp.RAW_SIZE)
    assert np.array_equal(rightf, oracle.remap_bilinear(src8.astype(np.float32), mx, my, 9))
    assert off == len(raw)


def test_dropin_kitti_io_round_trip(dropin, tmp_path):
    """load_gt_pose / save_txt / eval_pose / save_to_vis of run_odometry_kitti_offline.cpp:286-473 (host only)."""
    _, driver = dropin
    rng = np.random.default_rng(1)
    poses = rng.normal(size=(5, 12)).astype(np.float32) * 10
    with open(tmp_path / "in.txt", "w") as f:
        for p in poses:
            f.write(" ".join("%e" % v for v in p) + "\n")   # the devkit writes scientific notation
    r = subprocess.run([driver, "io", str(tmp_path / "in.txt"), "5", str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    back = np.loadtxt(tmp_path / "poses_out.txt", dtype=np.float64)
    lines = open(tmp_path / "poses_out.txt").read().splitlines()
    assert len(lines) == 5 and all(len(l.split(" ")) == 12 for l in lines)
    assert lines[0].split(" ")[0] == "%f" % np.float32(float("%e" % poses[0, 0]))  # std::to_string(float) == "%f"
    assert np.allclose(back, poses, atol=1e-6 + 1e-6 * np.abs(poses).max())
    err, n = open(tmp_path / "err.txt").read().split()
    assert float(err) == pytest.approx(0.5, abs=1e-5) and int(n) == 5
    y, x = np.mgrid[0:4, 0:8]
    g = open(tmp_path / "gray_3.pgm", "rb").read()
    assert g.startswith(b"P5\n8 4\n255\n")
    want = np.clip(np.rint((40.0 * x - 20.0 + 0.5 * y).astype(np.float32)), 0, 255).astype(np.uint8)
    assert np.array_equal(np.frombuffer(g[-32:], np.uint8).reshape(4, 8), want)
    m = np.frombuffer(open(tmp_path / "mask_3.pgm", "rb").read()[-32:], np.uint8).reshape(4, 8)
    assert np.array_equal(m, ((x + y) & 1).astype(np.uint8))
    d = np.frombuffer(open(tmp_path / "disparity_3.pgm", "rb").read()[-64:], ">u2").reshape(4, 8)
    dep = (np.float32(0.01) * (x + 1).astype(np.float32) * (y + 1).astype(np.float32)).astype(np.float32)
    assert np.array_equal(d, np.where(m != 0, (np.float32(386.1448) * dep).astype(np.uint16), 0))
    assert open(tmp_path / "keyframe_id.txt").read().split() == ["0", "7", "19"]


def test_dropin_stereo_rectify_matches_cv2_golden(dropin, tmp_path):
    """The drop-in's restatement of cv::stereoRectify (used by SetUpStereoCameraSystem when the build has no calib3d) on the
    reference's calibration file against cv2 4.13.0 (tests/golden/cv2_camera.npz) -- host arithmetic, no GPU."""
    _, driver = dropin
    _write_camchain(tmp_path / "camchain.yaml")
    G = np.load(os.path.join(ROOT, "tests", "golden", "cv2_camera.npz"))
    r = subprocess.run([driver, "stereo", str(tmp_path / "camchain.yaml"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "o.bin", "rb").read()
    d = np.frombuffer(raw, np.float64, 42)
    rois = np.frombuffer(raw, np.int32, 8, 42 * 8)
    R1, P1, R2, P2 = d[:9].reshape(3, 3), d[9:21].reshape(3, 4), d[21:30].reshape(3, 3), d[30:42].reshape(3, 4)
    assert np.abs(R1 - G["R1"]).max() < 1e-14 and np.abs(R2 - G["R2"]).max() < 1e-14
    assert np.abs(P1 - G["P1"]).max() < 1e-6 and np.abs(P2 - G["P2"]).max() < 1e-6
    assert np.array_equal(P1[:, 2], G["P1"][:, 2]) and np.array_equal(P2[:, 2], G["P2"][:, 2])   # principal points: exact
    assert tuple(rois[:4]) == tuple(G["roi1"]) and tuple(rois[4:]) == tuple(G["roi2"])
    # the rectified images do not depend on the 1e-9 relative difference of the focal length: same fixed-point tables
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import camera_patterns as cp
    for K, D, R, P, Rg, Pg in (("K0", "D0", R1, P1, "R1", "P1"), ("K1", "D1", R2, P2, "R2", "P2")):
        mx, my = oracle.init_undistort_rectify_map(G[K], G[D], R, P, cp.RAW_SIZE)
        gx, gy = oracle.init_undistort_rectify_map(G[K], G[D], G[Rg], G[Pg], cp.RAW_SIZE)
        fx_ = lambda m: np.rint(m * np.float32(32)).astype(np.int64)
        assert (fx_(mx) != fx_(gx)).mean() < 1e-4 and (fx_(my) != fx_(gy)).mean() < 1e-4
        assert np.abs(mx - gx).max() < 1e-4 and np.abs(my - gy).max() < 1e-4


@pytest.mark.parametrize("case", [0, 1, 2], ids=["converging", "right_camera_on_the_left", "vertical_pair"])
def test_dropin_stereo_rectify_other_geometries(dropin, tmp_path, case):
    """More stereo geometries against cv2 (strong convergence, negative baseline, baseline along y)."""
    _, driver = dropin
    G = np.load(os.path.join(ROOT, "tests", "golden", "cv2_camera.npz"))
    g = lambda k: G[f"stereo{case}/{k}"]
    np.concatenate([g("K0").ravel(), g("D0").ravel(), g("K1").ravel(), g("D1").ravel(), g("R").ravel(), g("t").ravel()]).astype(
        np.float64).tofile(tmp_path / "in.f64")
    r = subprocess.run([driver, "stereo_raw", str(tmp_path / "in.f64"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "o.bin", "rb").read()
    d = np.frombuffer(raw, np.float64, 42)
    rois = np.frombuffer(raw, np.int32, 8, 42 * 8)
    R1, P1, R2, P2 = d[:9].reshape(3, 3), d[9:21].reshape(3, 4), d[21:30].reshape(3, 3), d[30:42].reshape(3, 4)
    assert np.abs(R1 - g("R1")).max() < 1e-13 and np.abs(R2 - g("R2")).max() < 1e-13
    # OpenCV forms the scale of the new focal length from rectangles it keeps in float: agreement is to float precision
    assert np.allclose(P1, g("P1"), rtol=5e-7, atol=1e-9) and np.allclose(P2, g("P2"), rtol=5e-7, atol=1e-9)
    assert np.array_equal(P1[:, 2], g("P1")[:, 2]) and np.array_equal(P2[:, 2], g("P2")[:, 2])   # principal points: exact
    assert np.abs(rois[:4] - g("roi1")).max() <= 1 and np.abs(rois[4:] - g("roi2")).max() <= 1


@pytest.mark.gpu
def test_dropin_set_up_stereo_camera_system(dropin, tmp_path):
    """SetUpStereoCameraSystem end to end (test_camera_setup.cpp:19): baseline, valid region, rectified intrinsics."""
    _, driver = dropin
    _write_camchain(tmp_path / "camchain.yaml")
    G = np.load(os.path.join(ROOT, "tests", "golden", "cv2_camera.npz"))
    r = subprocess.run([driver, "setup", str(tmp_path / "camchain.yaml"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "stereo configuration done!" in r.stdout
    raw = open(tmp_path / "o.bin", "rb").read()
    baseline = np.frombuffer(raw, np.float64, 1)[0]
    valid = tuple(np.frombuffer(raw, np.int32, 4, 8))
    k = np.frombuffer(raw, np.float64, 8, 24).reshape(2, 4)
    P2 = G["P2"]
    assert baseline == pytest.approx(abs(P2[0, 3] / P2[0, 0]), rel=1e-9)
    r1, r2 = G["roi1"], G["roi2"]
    x0, y0 = max(r1[0], r2[0]), max(r1[1], r2[1])
    x1, y1 = min(r1[0] + r1[2], r2[0] + r2[2]), min(r1[1] + r1[3], r2[1] + r2[3])
    assert valid == (x0, y0, x1 - x0, y1 - y0)                       # src/camera.cpp:151-157
    for i, P in enumerate((G["P1"], G["P2"])):
        assert k[i] == pytest.approx([P[0, 0], P[1, 1], P[0, 2], P[1, 2]], rel=1e-8)
