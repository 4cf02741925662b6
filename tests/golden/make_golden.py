"""Generate the golden vectors of tests/golden/ with the in-container cv2 (4.13.0).

The reference calls cv::GaussianBlur / cv::pyrDown / cv::medianBlur (OpenCV, un-vendored); these
fixtures pin the oracle's restatement of those primitives to what cv2 computes here.  Run from the
repository root:  python tests/golden/make_golden.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


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


if __name__ == "__main__":
    main()
