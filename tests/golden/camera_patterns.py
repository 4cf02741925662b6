"""Inputs of the camera fixtures (numpy only, no cv2): the calibration values of the reference's
calibration_file/camchain.yaml and deterministic raw frames / stress maps.  Shared by make_golden.py (which
runs cv2 on them) and the tests (which run the oracle and the CUDA path on them)."""
import hashlib

import numpy as np

# the values of the reference's calibration_file/camchain.yaml (cam0 / cam1 / T_cn_cnm1)
CAM0_INTR = [427.32814323885566, 429.48081105226316, 367.1148716890002, 242.03387791215218]
CAM0_DIST = [-0.35292630520315216, 0.09970701156068408, -0.0003265055193558261, -0.003400767380536901]
CAM1_INTR = [425.28226969376584, 427.5013362691404, 342.6156277602674, 233.38645927695092]
CAM1_DIST = [-0.34242635946786465, 0.09353275937137827, 0.000332922660566574, -0.001440982693394223]
T_CN_CNM1 = [[0.9999842188801975, 0.0051761908815188395, -0.0021838128387943914, -0.060400809282521006],
             [-0.005183031408333682, 0.9999816430410434, -0.0031384337416914382, 0.00020747637203608188],
             [0.002167527618515751, 0.0031497029842259693, 0.9999926905708509, 3.97878900435667e-05]]
RAW_SIZE = (640, 480)  # (width, height): src/camera.cpp:75,100


def kmat(i):
    return np.array([[i[0], 0, i[2]], [0, i[1], i[3]], [0, 0, 1]], np.float64)


def raw_pattern_u8(w=640, h=480):
    """deterministic integer texture (no libm): the raw camera frame of the remap fixtures"""
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    return ((x * x * 3 + y * 7 + (x * y) // 5 + (y * y) // 3) % 256).astype(np.uint8)


def raw_pattern_f32(w=640, h=480):
    y, x = np.mgrid[0:h, 0:w].astype(np.int64)
    return (raw_pattern_u8(w, h).astype(np.float32) + ((x * 13 + y * 29) % 64).astype(np.float32) / np.float32(64))


def stress_maps(w=640, h=480):
    rng = np.random.default_rng(20190117)
    mx = (rng.random((h, w), dtype=np.float32) * np.float32(700) - np.float32(30)).astype(np.float32)
    my = (rng.random((h, w), dtype=np.float32) * np.float32(540) - np.float32(30)).astype(np.float32)
    return mx, my


def digest(a):
    return np.array(hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest())
