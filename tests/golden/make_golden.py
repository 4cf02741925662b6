"""Generate the golden vectors of tests/golden/ with the in-container cv2 (4.13.0).

The reference calls cv::GaussianBlur / cv::pyrDown / cv::medianBlur (OpenCV, un-vendored); these
fixtures pin the oracle's restatement of those primitives to what cv2 computes here.  cv2_camera.npz does
the same for the undistort + rectify path (src/camera.cpp:40-83 and :84-170 of the reference:
cv::stereoRectify, cv::initUndistortRectifyMap, cv::remap) on the reference's own calibration file.
Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main():
    rng = np.random.default_rng(20190113)
    out = {"cv2_version": np.array(cv2.__version__)}
    cases = {"u8_157x93": rng.integers(0, 256, (93, 157)).astype(np.float32),
             "f32_157x93": (rng.random((93, 157)) * 255).astype(np.float32),
             "u8_64x48": rng.integers(0, 256, (48, 64)).astype(np.float32)}
    for name, img in cases.items():
        out[f"{name}/img"] = img
        out[f"{name}/blur"] = cv2.GaussianBlur(img, (3, 3), 0)
        out[f"{name}/median"] = cv2.medianBlur(img, 3)
        cur = img
        for l in range(1, 4):
            dw, dh = cur.shape[1] // 2, cur.shape[0] // 2
            cur = cv2.pyrDown(cur, dstsize=(dw, dh))
            out[f"{name}/down{l}"] = cur
    np.savez_compressed(os.path.join(HERE, "cv2_primitives.npz"), **out)
    print("wrote", os.path.join(HERE, "cv2_primitives.npz"))


from camera_patterns import (CAM0_DIST, CAM0_INTR, CAM1_DIST, CAM1_INTR, RAW_SIZE, T_CN_CNM1, digest, kmat,  # noqa: E402
                             raw_pattern_f32, raw_pattern_u8, stress_maps)


def camera():
    K0, K1 = kmat(CAM0_INTR), kmat(CAM1_INTR)
    D0, D1 = np.array(CAM0_DIST), np.array(CAM1_DIST)
    T = np.array(T_CN_CNM1)
    R, t = T[:, :3].copy(), T[:, 3:].copy()
    out = {"cv2_version": np.array(cv2.__version__), "K0": K0, "K1": K1, "D0": D0, "D1": D1, "R": R, "t": t}
    # SetUpStereoCameraSystem, src/camera.cpp:139-141
    R1, R2, P1, P2, Q, roi1, roi2 = cv2.stereoRectify(K0, D0, K1, D1, RAW_SIZE, R, t, flags=cv2.CALIB_ZERO_DISPARITY,
                                                      alpha=1, newImageSize=RAW_SIZE)
    out.update(R1=R1, R2=R2, P1=P1, P2=P2, Q=Q, roi1=np.array(roi1), roi2=np.array(roi2))
    src8, srcf = raw_pattern_u8(), raw_pattern_f32()
    out["src8_sha"], out["srcf_sha"] = digest(src8), digest(srcf)
    for name, K, D, Rr, P in (("left", K0, D0, R1, P1), ("right", K1, D1, R2, P2)):
        # ConfigureCamera, src/camera.cpp:69
        mx, my = cv2.initUndistortRectifyMap(K, D, Rr, P, RAW_SIZE, cv2.CV_32FC1)
        out[f"{name}/mapx_sha"], out[f"{name}/mapy_sha"] = digest(mx), digest(my)
        out[f"{name}/mapx_rows"], out[f"{name}/mapy_rows"] = mx[::16].copy(), my[::16].copy()
        # UndistortRectify with its defaults, src/camera.cpp:72,81
        for tag, src in (("u8", src8), ("f32", srcf)):
            for cval in (0, 9):
                d = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=cval)
                out[f"{name}/remap_{tag}_c{cval}_sha"] = digest(d)
                out[f"{name}/remap_{tag}_c{cval}_rows"] = d[::16].copy()
    # maps that leave the image on every side, land on the last row / column and on exact texel centres
    mx, my = stress_maps()
    out["stress/maps_sha"] = digest(np.stack([mx, my]))
    for tag, src in (("u8", src8), ("f32", srcf)):
        d = cv2.remap(src, mx, my, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=9)
        out[f"stress/remap_{tag}_sha"] = digest(d)
        out[f"stress/remap_{tag}_rows"] = d[::16].copy()
    # more stereo geometries for the drop-in's stereoRectify restatement (dropin/src/stereo_rectify.cpp): a strongly
    # converging pair, a pair with the right camera on the other side, and a vertical pair (baseline along y)
    rng = np.random.default_rng(20190124)
    for i, (rvec, tvec) in enumerate([((0.03, -0.08, 0.02), (-0.12, 0.004, 0.01)), ((-0.01, 0.05, -0.03), (0.09, -0.002, -0.004)),
                                      ((0.02, 0.01, -0.015), (0.003, -0.11, 0.006))]):
        Ka = kmat([420 + 20 * rng.random(), 425 + 20 * rng.random(), 320 + 30 * rng.random(), 240 + 20 * rng.random()])
        Kb = kmat([420 + 20 * rng.random(), 425 + 20 * rng.random(), 320 + 30 * rng.random(), 240 + 20 * rng.random()])
        Da = np.array([-0.3 + 0.1 * rng.random(), 0.08 + 0.04 * rng.random(), 1e-3 * (rng.random() - 0.5), 1e-3 * (rng.random() - 0.5)])
        Db = np.array([-0.3 + 0.1 * rng.random(), 0.08 + 0.04 * rng.random(), 1e-3 * (rng.random() - 0.5), 1e-3 * (rng.random() - 0.5)])
        Rm = cv2.Rodrigues(np.array(rvec, np.float64))[0]
        tv = np.array(tvec, np.float64).reshape(3, 1)
        r1, r2, p1, p2, _, ro1, ro2 = cv2.stereoRectify(Ka, Da, Kb, Db, RAW_SIZE, Rm, tv, flags=cv2.CALIB_ZERO_DISPARITY, alpha=1,
                                                       newImageSize=RAW_SIZE)
        out.update({f"stereo{i}/K0": Ka, f"stereo{i}/K1": Kb, f"stereo{i}/D0": Da, f"stereo{i}/D1": Db, f"stereo{i}/R": Rm,
                    f"stereo{i}/t": tv, f"stereo{i}/R1": r1, f"stereo{i}/R2": r2, f"stereo{i}/P1": p1, f"stereo{i}/P2": p2,
                    f"stereo{i}/roi1": np.array(ro1), f"stereo{i}/roi2": np.array(ro2)})
    np.savez_compressed(os.path.join(HERE, "cv2_camera.npz"), **out)
    print("wrote", os.path.join(HERE, "cv2_camera.npz"))


if __name__ == "__main__":
    main()
    camera()
