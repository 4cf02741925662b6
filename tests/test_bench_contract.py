"""bench.py --impl reference runs without a GPU: its JSON line must carry the contract's keys for every workload
(the GPU arm prints the same line plus roofline / clocks, checked on the GPU box by the driver)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run_reference(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *extra], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    return json.loads(lines[0])


@pytest.mark.parametrize("workload,unit", [("kitti", "frames/s"), ("rectify", "pairs/s")])
def test_reference_arm_line(workload, unit):
    d = run_reference("--workload", workload, "--steps", "1", "--warmup", "0", "--seqs", "3")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["unit"] == unit and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"]
    if workload == "kitti":
        # the reference arm runs the SAME number of sequences per step as the GPU arm's config (same_config)
        assert d["config"]["sequences_per_gpu"] == 3 and d["frames_per_step"] == 3
        assert d["pose_error_vs_gt"]["max"] < 0.05      # one 0.8 m frame from the keyframe: centimetres
    cb = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["kind"] in ("reference", "port")
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and "host:" in cb["sample"]
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == unit and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    """under torchrun only rank 0 prints the reference line; the other ranks exit 0 without work"""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)


def test_bench_host_helpers():
    """host logic of bench.py that needs no GPU: weak / strong sequence split, ground-truth poses, pose-error summary,
    sysfs cpulist parsing of the NUMA placement step"""
    import argparse
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench

    a = argparse.Namespace(seqs=64, seqs_total=0)
    assert bench.seqs_per_gpu(a, 8) == (64, "weak")
    a.seqs_total = 64
    assert bench.seqs_per_gpu(a, 8) == (8, "strong") and bench.seqs_per_gpu(a, 1) == (64, "strong")
    with pytest.raises(SystemExit):
        bench.seqs_per_gpu(argparse.Namespace(seqs=64, seqs_total=10), 4)
    gt = bench.gt_relative_poses(6)
    assert np.allclose(gt[0], np.eye(4)) and abs(gt[5][2, 3] - 4.0) < 1e-3        # 0.8 m per frame forward
    est = np.stack([gt[5], gt[5]]).astype(np.float32)
    est[1, 0, 3] += 0.3
    s = bench.pose_error_summary(est, gt[5])
    assert s["max"] == pytest.approx(0.3, abs=1e-6) and s["median"] == pytest.approx(0.15, abs=1e-6)
    assert s["max_rel"] == pytest.approx(0.3 / s["path_length_m"], rel=1e-6)
    assert bench._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}


def test_stdout_is_one_json_line_even_if_a_library_prints():
    """bench.py points fd 1 at stderr and writes its JSON line to the original stdout: anything a library prints to stdout
    (NCCL's INFO log) cannot break the one-line contract"""
    code = ("import sys, os; sys.path.insert(0, %r); import bench; bench.stdout_to_stderr(); "
            "os.write(1, b'NCCL INFO noise\\n'); print('more noise'); bench.emit({'ok': 1})" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert r.returncode == 0 and r.stdout == '{"ok": 1}\n' and "NCCL INFO noise" in r.stderr and "more noise" in r.stderr
