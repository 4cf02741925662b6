import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _device():
    import torch
    return "cuda" if torch.cuda.is_available() else "cpu"


class SynthData:
    """Synthetic KITTI-shaped frames shared by the tests (rendered once per session)."""

    def __init__(self):
        self._cache = {}

    def pair(self, seed_id=0, w=1241, h=376):
        """Config-1 pair: frame 1 stereo + frame 2 left after the known motion of SURVEY 8(d)."""
        key = ("pair", seed_id, w, h)
        if key not in self._cache:
            from odometry_b200 import synth
            sc = synth.make_scene(seed_id, 2)
            T1 = synth.pose_from_motion()
            T2 = synth.pose_from_motion(0.02, -0.01, 0.30, math.radians(0.2), math.radians(0.5), math.radians(-0.1))
            dev = _device()
            L1, R1, z1 = synth.render_stereo(sc, T1, w, h, dev)
            L2, _ = synth.render(sc, T2, w, h, dev)
            self._cache[key] = dict(
                L1=L1.cpu().numpy().astype(np.float32), R1=R1.cpu().numpy().astype(np.float32),
                L2=L2.cpu().numpy().astype(np.float32), z1=z1.cpu().numpy(),
                T_gt=synth.relative_pose(T1, T2))
        return self._cache[key]

    def sequence(self, seed_id=0, n_frames=6, w=1241, h=376, step=0.8):
        key = ("seq", seed_id, n_frames, w, h, step)
        if key not in self._cache:
            from odometry_b200 import synth
            sc = synth.make_scene(seed_id, n_frames, step)
            poses = synth.trajectory(n_frames, step=step)
            dev = _device()
            frames = []
            for T in poses:
                L, R, _ = synth.render_stereo(sc, T, w, h, dev)
                frames.append((L.cpu().numpy(), R.cpu().numpy()))
            self._cache[key] = dict(frames=frames, poses=poses)
        return self._cache[key]


@pytest.fixture(scope="session")
def synth_data():
    return SynthData()


@pytest.fixture(scope="session")
def ctx():
    from odometry_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
