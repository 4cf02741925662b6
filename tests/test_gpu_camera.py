"""CUDA undistort + rectify (odo_camera_*, camera.cu) against the oracle and the cv2 golden vectors -- through the
C-ABI.  Reference: CameraPyramid::ConfigureCamera / UndistortRectify, src/camera.cpp:40-83."""
import os
import sys

import numpy as np
import pytest

import oracle
from odometry_b200 import capi

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import camera_patterns as cp  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "cv2_camera.npz"))
SIDES = [("left", "K0", "D0", "R1", "P1"), ("right", "K1", "D1", "R2", "P2")]


def sha(a):
    return str(cp.digest(a))


def make_camera(ctx, side, levels=4):
    name, K, D, R, P = side
    cam = capi.Camera(ctx, levels, GOLD[K], GOLD[D], sensor=(5.76, 4.29), resolution=(640, 482))
    cam.configure(GOLD[R], GOLD[P], cp.RAW_SIZE)
    return cam


@pytest.mark.parametrize("side", SIDES, ids=["left", "right"])
def test_configure_maps_and_intrinsics(ctx, side):
    cam = make_camera(ctx, side)
    mx, my = cam.maps()
    ox, oy = oracle.init_undistort_rectify_map(GOLD[side[1]], GOLD[side[2]], GOLD[side[3]], GOLD[side[4]], cp.RAW_SIZE)
    assert np.array_equal(mx.view(np.uint32), ox.view(np.uint32))
    assert np.array_equal(my.view(np.uint32), oy.view(np.uint32))
    assert sha(mx) == str(GOLD[f"{side[0]}/mapx_sha"])  # cv2 itself (the y map of the right eye: see test_oracle_camera)
    ok = oracle.camera_intrinsic_pyramid(GOLD[side[4]], 4)
    for l in range(4):
        assert np.array_equal(cam.intrinsics(l), ok[l])
    k = cam.lm_intrinsics()
    assert (k.fx, k.cx, k.cy) == (np.float32(ok[0, 0]), np.float32(ok[0, 2]), np.float32(ok[0, 3]))
    with pytest.raises(capi.OdoError):
        cam.intrinsics(4)
    cam.close()


@pytest.mark.parametrize("side", SIDES, ids=["left", "right"])
@pytest.mark.parametrize("cval", [0, 9])
def test_undistort_rectify_matches_oracle_and_cv2_golden(ctx, side, cval):
    cam = make_camera(ctx, side)
    mx, my = cam.maps()
    src8, srcf = cp.raw_pattern_u8(), cp.raw_pattern_f32()
    d8 = cam.undistort_rectify(src8, border_value=cval)
    assert d8.dtype == np.uint8 and np.array_equal(d8, oracle.remap_bilinear(src8, mx, my, cval))
    assert sha(d8) == str(GOLD[f"{side[0]}/remap_u8_c{cval}_sha"])
    df = cam.undistort_rectify(srcf, border_value=cval)
    assert np.array_equal(df.view(np.uint32), oracle.remap_bilinear(srcf, mx, my, cval).view(np.uint32))
    assert sha(df) == str(GOLD[f"{side[0]}/remap_f32_c{cval}_sha"])
    # fused convertTo(CV_32F): the float image of a u8 frame is the u8 result, exactly
    d8f = cam.undistort_rectify(src8, out_dtype=np.float32, border_value=cval)
    assert np.array_equal(d8f, d8.astype(np.float32))
    cam.close()


def test_stress_maps_leave_the_frame_on_every_side(ctx):
    cam = capi.Camera(ctx, 1, np.eye(3), np.zeros(4))
    mx, my = cp.stress_maps()
    cam.set_maps(mx, my)
    for tag, src in (("u8", cp.raw_pattern_u8()), ("f32", cp.raw_pattern_f32())):
        d = cam.undistort_rectify(src, border_value=9)
        assert np.array_equal(d, oracle.remap_bilinear(src, mx, my, 9))
        assert sha(d) == str(GOLD[f"stress/remap_{tag}_sha"])
    cam.close()


@pytest.mark.parametrize("dw,dh,sw,sh", [(333, 100, 83, 61), (5, 3, 2, 2), (1241, 376, 1300, 400), (64, 1, 640, 480)])
def test_ragged_sizes_and_special_coordinates(ctx, dw, dh, sw, sh):
    """widths that are not a multiple of the 4-pixel vector, tiny sources, exact texel centres, the last row and
    column, NaN / huge coordinates"""
    rng = np.random.default_rng(dw * 1000 + dh)
    mx = (rng.random((dh, dw), dtype=np.float32) * (sw + 8) - 4).astype(np.float32)
    my = (rng.random((dh, dw), dtype=np.float32) * (sh + 8) - 4).astype(np.float32)
    flat_x, flat_y = mx.ravel(), my.ravel()
    special = [(0.0, 0.0), (sw - 1.0, sh - 1.0), (sw - 1.0 + 1 / 64, 0.5), (-1.0, -1.0), (-0.015625, 1.0),
               (np.nan, 1.0), (1.0, np.inf), (-3e9, 3e9), (sw - 0.984375, sh - 0.984375), (0.015624, 0.015626)]
    for i, (a, b) in enumerate(special[:flat_x.size]):
        flat_x[i], flat_y[i] = a, b
    cam = capi.Camera(ctx, 1, np.eye(3), np.zeros(4))
    cam.set_maps(mx, my)
    s8 = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    sf = (rng.random((sh, sw)) * 255).astype(np.float32)
    assert np.array_equal(cam.undistort_rectify(s8, border_value=200), oracle.remap_bilinear(s8, mx, my, 200))
    got, ref = cam.undistort_rectify(sf, border_value=1.5), oracle.remap_bilinear(sf, mx, my, 1.5)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    cam.close()


def test_batched_device_variant(ctx):
    import torch
    cam = make_camera(ctx, SIDES[0])
    mx, my = cam.maps()
    rng = np.random.default_rng(3)
    B = 3
    src = rng.integers(0, 256, (B, 480, 640), dtype=np.uint8)
    ds = torch.from_numpy(src).cuda()
    dst_f = torch.empty((B, 480, 640), dtype=torch.float32, device="cuda")
    dst_8 = torch.empty((B, 480, 640), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    cam.undistort_rectify_dev(ds.data_ptr(), True, 640, 480, 640 * 480, dst_f.data_ptr(), False, 640 * 480, B)
    cam.undistort_rectify_dev(ds.data_ptr(), True, 640, 480, 640 * 480, dst_8.data_ptr(), True, 640 * 480, B)
    ctx.sync()
    for b in range(B):
        ref = oracle.remap_bilinear(src[b], mx, my, 0)
        assert np.array_equal(dst_8[b].cpu().numpy(), ref)
        assert np.array_equal(dst_f[b].cpu().numpy(), ref.astype(np.float32))
    cam.close()


def test_unconfigured_camera_and_bad_types_fail(ctx):
    cam = capi.Camera(ctx, 4, np.eye(3), np.zeros(4))
    with pytest.raises(capi.OdoError):
        cam.undistort_rectify(np.zeros((4, 4), np.uint8))
    with pytest.raises(capi.OdoError):
        cam.configure(np.zeros((3, 3)), np.zeros((3, 4)), (8, 8))  # singular
    cam.set_maps(np.zeros((4, 4), np.float32), np.zeros((4, 4), np.float32))
    with pytest.raises(capi.OdoError):
        cam.undistort_rectify(np.zeros((4, 4), np.float32), out_dtype=np.uint8)
    cam.close()


def test_tracker_raw_ingest_equals_rectify_then_track(ctx, synth_data):
    """odo_tracker_step_raw_u8 (remap fused in front of the frame's kernels) gives bit-identical poses to
    UndistortRectify on every frame followed by odo_tracker_step_u8."""
    w, h = 1241, 376
    seq = synth_data.sequence(0, 4)
    raw = [(np.clip(np.rint(L), 0, 255).astype(np.uint8), np.clip(np.rint(R), 0, 255).astype(np.uint8)) for L, R in seq["frames"]]
    K = np.array([[718.856, 0, 607.1928], [0, 718.856, 185.2157], [0, 0, 1.0]])
    P = np.hstack([K, np.zeros((3, 1))])
    cams = []
    for dist in ([-0.002, 0.0004, 1e-5, -2e-5], [-0.0018, 0.0003, -1e-5, 1e-5]):
        cam = capi.Camera(ctx, 4, K, dist)
        cam.configure(np.eye(3), P, (w, h))
        cams.append(cam)
    ta = capi.Tracker(ctx, w, h, 4, 1)
    tb = capi.Tracker(ctx, w, h, 4, 1)
    tb.set_rectify(cams[0], cams[1], w, h)
    for L8, R8 in raw:
        ra = ta.step([cams[0].undistort_rectify(L8)], [cams[1].undistort_rectify(R8)])
        rb = tb.step_raw_u8([L8], [R8])
        assert ra[0] == rb[0]
        assert np.array_equal(ra[1], rb[1]) and np.array_equal(ra[2], rb[2]) and np.array_equal(ra[3], rb[3]) and np.array_equal(ra[4], rb[4])
    small = make_camera(ctx, SIDES[0])  # 640 x 480 table: not the tracker's image size
    with pytest.raises(capi.OdoError):
        tb.set_rectify(cams[0], small, w, h)
    tc = capi.Tracker(ctx, w, h, 4, 1)
    with pytest.raises(capi.OdoError):
        tc.enqueue_raw_ptrs([raw[0][0].ctypes.data], [raw[0][1].ctypes.data])  # set_rectify has not run
    tc.close(); small.close()
    ta.close(); tb.close()
    for c in cams:
        c.close()


@pytest.mark.parametrize("n", [1, 3, 4, 9])
def test_batched_host_entry_and_image_groups(ctx, n):
    """odo_camera_undistort_rectify_batch: n host frames in, n out with one launch; the kernel shares a table entry
    between groups of four images, so batch sizes on both sides of a multiple of four are covered."""
    cam = make_camera(ctx, SIDES[1])
    mx, my = cam.maps()
    rng = np.random.default_rng(40 + n)
    src = rng.integers(0, 256, (n, 480, 640), dtype=np.uint8)
    out_f = np.zeros((n, 480, 640), np.float32)
    cam.undistort_rectify_batch_ptrs([src[i].ctypes.data for i in range(n)], True, 640, 480,
                                     [out_f[i].ctypes.data for i in range(n)], False, border_value=7)
    srcf = (src.astype(np.float32) + 0.25)
    out_ff = np.zeros((n, 480, 640), np.float32)
    cam.undistort_rectify_batch_ptrs([srcf[i].ctypes.data for i in range(n)], False, 640, 480,
                                     [out_ff[i].ctypes.data for i in range(n)], False, border_value=7)
    for i in range(n):
        assert np.array_equal(out_f[i], oracle.remap_bilinear(src[i], mx, my, 7).astype(np.float32)), i
        assert np.array_equal(out_ff[i].view(np.uint32), oracle.remap_bilinear(srcf[i], mx, my, 7).view(np.uint32)), i
    cam.close()
