#!/usr/bin/env python
"""bench.py -- KITTI-shape tracking throughput of the B200 hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: one new 1241x376 stereo frame for each of S
independent synthetic sequences resident on the GPU (image pyramid + LM Solve against the keyframe
+ ComputeDepth (selection, SSD disparity search, depth refinement) + depth pyramid + keyframe rule:
the per-frame body of run_odometry_kitti_offline.cpp:198-271).  With --gpus N every rank runs its
own S sequences (sharding of independent sequences, no data-path collective => weak scaling);
--seqs-total T shards T sequences over the ranks instead (config 4 as written: strong scaling).

  value : frames/s with the step's inputs (u8 frames) already resident in HBM (odo_tracker_enqueue_dev_u8)
  e2e   : frames/s through the reference-facing C-ABI call with HOST buffers: u8 images in pinned
          memory (what the reference's imread leaves, run_odometry_kitti_offline.cpp:334-359), H2D of
          the step's frames and D2H of the poses inside the timed region; the CV_32F variant of the
          same call is the sub-key e2e.f32_ingest

--impl reference times the reference's own CPU implementation (oracle/_ref: its hot-path sources compiled
against stand-in Eigen/OpenCV headers; the oracle port if that library is absent) on all host cores, one
sequence per thread.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, LEVELS = 1241, 376, 4
METRIC = "KITTI-shape tracking frames/s (pyramid + disparity/depth + LM align per frame)"
def host_cpu():
    """model name and logical core count of the host the CPU arm runs on (reported with the baseline)"""
    model = "unknown CPU"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return f"{model}, {os.cpu_count()} logical cores"


CPU_KIND_NOTE = {
    "reference": "the reference's own classes (oracle/_ref: its hot-path sources compiled unmodified against stand-in "
                 "Eigen/OpenCV/Sophus headers, the loop of run_odometry_kitti_offline.cpp:95-271 around them; the stand-in's "
                 "matrix products are scalar loops, so a build against real, vectorised Eigen may be faster); host: " + host_cpu(),
    "port": "oracle port of the reference (oracle/_ref not present); host: " + host_cpu(),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--seqs", type=int, default=64, help="independent sequences per GPU (batch of one step)")
    ap.add_argument("--seqs-total", type=int, default=0,
                    help="config 4 as written: this many sequences IN TOTAL, sharded over the ranks (strong scaling); "
                         "overrides --seqs")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-sequence latency and the 4K (config 5) legs")
    ap.add_argument("--scene", default="street", choices=["street", "isotropic"],
                    help="texture spectrum of the synthetic scene (odometry_b200/synth.py); isotropic = the round-1 scene on "
                         "which the reference's own loop loses track at the first keyframe switch")
    ap.add_argument("--streams", type=int, default=1, help="independent trackers (contexts/streams) per GPU sharing --seqs")
    ap.add_argument("--cpu-seqs", type=int, default=2, help="sequences of the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workload", default="kitti", choices=["kitti", "4k", "rectify"],
                    help="kitti: the headline tracking step (BASELINE.json configs 3/4); 4k: config 5, 3840x2160 "
                         "6-level pyramid + all-pixel dense LM alignment of each new frame against a keyframe; "
                         "rectify: the camera front end (SURVEY 8 f3), UndistortRectify of raw 640x480 stereo pairs")
    return ap.parse_args()


def seqs_per_gpu(args, world):
    """(S, scaling): --seqs per GPU (weak) or --seqs-total sharded over the ranks (strong, config 4 as written)."""
    if args.seqs_total > 0:
        if args.seqs_total % world:
            raise SystemExit(f"--seqs-total {args.seqs_total} does not divide over {world} ranks")
        return args.seqs_total // world, "strong"
    return args.seqs, "weak"


def config_dict(args, world=None, extra=None):
    world = world or args.gpus
    S, scaling = seqs_per_gpu(args, world)
    cfg = {
        "workload": f"config 4/3: {S} independent synthetic {W}x{H} stereo sequences per GPU"
                    + (f" ({args.seqs_total} in total, sharded)" if scaling == "strong" else "")
                    + f", {LEVELS}-level pyramid, full tracking step per frame (reference-faithful: depth every frame)",
        "sequences_per_gpu": S, "frame": f"{W}x{H}", "levels": LEVELS,
        "lm": "lambda 0.01, precision 0.995, iters {10,20,30,30}, Huber 28",
        "depth": "grad_th 8, ssd_th 900, photo_th 15, 0.1-30 m, 50 iters, boundary 4",
        "trajectory": "0.8 m/frame forward, yaw 0.4deg*sin(2*pi*t/100), seeds 20190113+sequence_id",
        "scene": {"street": "street corridor with the anisotropic 'street' spectrum (odometry_b200/synth.py STREET_BANDS: stripes "
                            "along the driving direction + long-wave structure; documented deviation from SURVEY 8(d), on which "
                            "the reference's own loop keeps track across keyframe switches)",
                  "isotropic": "street corridor with the round-1 isotropic spectrum (the reference's loop loses track on it)"}[args.scene],
        "parallelism": f"sequence sharding x{world} GPUs x{args.streams} trackers per GPU, no collective",
        "l2": "each step reads a fresh batch of frames (S*3.7 MB > 126 MB L2 for S >= 34); no explicit flush",
    }
    if extra:
        cfg.update(extra)
    return cfg


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if self.nv and self.samples:
            return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": float(self.max_mhz), "reasons": sorted(self.reasons)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


_REAL_STDOUT = None


def stdout_to_stderr():
    """Everything any library writes to stdout from here on (NCCL's INFO log when NCCL_DEBUG=INFO is in the environment:
    it cannot be re-routed from inside the process, tools/nccl_dbg_test.py) lands on stderr; the one JSON line goes to the
    ORIGINAL stdout through emit().  bench.py does not touch NCCL_DEBUG."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def init_dist(dev, world):
    """NCCL only carries the barrier / max-over-ranks plumbing around the timed region."""
    if world <= 1:
        return
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)


def _parse_cpulist(txt):
    cpus = set()
    for part in txt.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa(local_rank):
    """Host placement for the e2e legs: run this process on the cores of the NUMA node the GPU hangs off BEFORE the pinned
    frame buffers are allocated (first touch puts them in that node's memory), so that with 8 ranks the host->device
    copies do not all cross one socket (every rank sat on cores 0-31 / node 0)."""
    info = {"numa_node": None, "cpus": None, "bound": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        info["pci"] = bdf
        info["numa_node"] = node
        cpus = _parse_cpulist(open(base + "/local_cpulist").read())
        allowed = os.sched_getaffinity(0)
        if node >= 0 and cpus & allowed:
            os.sched_setaffinity(0, cpus & allowed)
            info["bound"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as e:   # placement is best effort (containers without sysfs topology)
        info["note"] = f"not bound: {type(e).__name__}: {e}"
    return info


def h2d_probe_gbs(dev, stream, mbytes=256):
    """pinned host -> device copy rate of this rank alone (what bounds the e2e legs)"""
    import torch
    host = torch.empty(mbytes << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mbytes << 20, dtype=torch.uint8, device=dev)
    best = 0.0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        d.copy_(host, non_blocking=True)
        e1.record(stream)
        e1.synchronize()
        best = max(best, (mbytes << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    return hbm_peak, peak_src


def render_sequences(seq_ids, n_frames, device, scene="street"):
    """u8 frames [n_frames][len(seq_ids)][2][H][W] on `device` (torch) of the given global sequence ids."""
    import torch
    from odometry_b200 import synth
    n_seq = len(seq_ids)
    out = torch.empty((n_frames, n_seq, 2, H, W), dtype=torch.uint8, device=device)
    poses = synth.trajectory(n_frames, step=0.8)
    for s, sid in enumerate(seq_ids):
        sc = synth.make_scene(sid, n_frames, 0.8, bands=synth.ISOTROPIC_BANDS if scene == "isotropic" else synth.STREET_BANDS)
        for f, T in enumerate(poses):
            L, R, _ = synth.render_stereo(sc, T, W, H, device)
            out[f, s, 0], out[f, s, 1] = L, R
    return out


def gt_relative_poses(n_frames):
    """ground-truth pose of frame f in the frame-0 camera (what the tracker's pose_abs estimates)"""
    from odometry_b200 import synth
    poses = synth.trajectory(n_frames, step=0.8)
    P0i = np.linalg.inv(poses[0])
    return [P0i @ P for P in poses]


def pose_error_summary(pose_abs_by_seq, gt):
    """translation error of the estimated absolute poses ([S][4][4], row/col) against the ground truth 4x4"""
    err = np.linalg.norm(np.asarray(pose_abs_by_seq)[:, :3, 3] - gt[:3, 3][None, :], axis=1)
    dist = float(np.linalg.norm(gt[:3, 3]))
    return {"unit": "m", "path_length_m": dist, "median": float(np.median(err)), "max": float(err.max()),
            "median_rel": float(np.median(err) / max(dist, 1e-9)), "max_rel": float(err.max() / max(dist, 1e-9))}


def cpu_tracker_factory():
    """(kind, make_tracker): the reference's own classes (oracle/_ref: its first-party sources compiled against
    stand-in Eigen/OpenCV headers) when that library exists, else the oracle port.  Both in the reference's own
    arithmetic (fp32 sequential sums)."""
    import oracle
    from oracle import ref
    lp = lambda: oracle.lm_params(sum_mode=oracle.SUM_FLOAT_SEQ)
    dp = lambda: oracle.depth_params(sum_mode=oracle.SUM_FLOAT_SEQ)
    if ref.available():
        try:
            ref.lib()
            return "reference", lambda: ref.Tracker(W, H, LEVELS, lp(), dp())
        except OSError:
            pass
    oracle.lib()
    return "port", lambda: oracle.Tracker(W, H, LEVELS, lp(), dp())


def cpu_tracking_fps(frames_u8, n_seq, n_frames, threads):
    """CPU tracking loop (cpu_tracker_factory): n_seq sequences over `threads` host threads; frame 0
    (keyframe init) is untimed like the GPU arm's. Returns (frames/s, timed frames, kind)."""
    kind, make = cpu_tracker_factory()
    trackers = [make() for _ in range(n_seq)]
    fl = frames_u8.astype(np.float32)
    for s, t in enumerate(trackers):
        t.step(fl[0, s, 0], fl[0, s, 1])

    def work(s):
        for f in range(1, n_frames):
            trackers[s].step(fl[f, s, 0], fl[f, s, 1])

    t0 = time.perf_counter()
    if threads <= 1:
        for s in range(n_seq):
            work(s)
    else:
        ths = [threading.Thread(target=work, args=(s,)) for s in range(n_seq)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    dt = time.perf_counter() - t0
    for t in trackers:
        t.close()
    return n_seq * (n_frames - 1) / dt, n_seq * (n_frames - 1), kind


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU code on all host cores: the SAME sequences-per-step as the GPU arm
    (config), spread round robin over one worker thread per logical core (the ctypes calls release the GIL)."""
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    S, scaling = seqs_per_gpu(args, max(world, 1))
    n_seq = S * max(world, 1)
    n_frames = 1 + args.warmup + args.steps
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    frames = render_sequences(list(range(n_seq)), n_frames, dev, args.scene).cpu().numpy()
    kind, make = cpu_tracker_factory()
    trackers = [make() for _ in range(n_seq)]
    threads = max(1, min(cores, n_seq))
    last = [None] * n_seq
    n_kf = [0] * n_seq

    def advance(k, f0, f1):
        for s in range(k, n_seq, threads):
            for f in range(f0, f1):
                last[s] = trackers[s].step(frames[f, s, 0].astype(np.float32), frames[f, s, 1].astype(np.float32))
                n_kf[s] += int(bool(last[s][3])) if f > args.warmup else 0

    def parallel(f0, f1):
        ths = [threading.Thread(target=advance, args=(k, f0, f1)) for k in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()

    # warm-up steps are part of the same trajectory: time only the last `steps` frames
    parallel(0, 1 + args.warmup)
    t0 = time.perf_counter()
    parallel(1 + args.warmup, n_frames)
    dt = time.perf_counter() - t0
    fps = n_seq * args.steps / dt
    gt = gt_relative_poses(n_frames)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, max(world, 1)),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                         "sample": f"{n_seq} sequences x {args.steps} frames over {threads} host threads (round robin); "
                                   + CPU_KIND_NOTE[kind]},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "frames_per_step": n_seq,
        "pose_error_vs_gt": pose_error_summary([r[1] for r in last], gt[n_frames - 1]),
        "keyframes_in_timed_region": int(sum(n_kf)),
    }
    emit(line)


# ---------------------------------------------------------------------------------------------------------
# config 5: 3840x2160, 6 levels, dense alignment (the HBM-bound stress case; no ComputeDepth: the reference
# locks it to 376x1241, src/depth_estimate.cpp:46)
# ---------------------------------------------------------------------------------------------------------
W4, H4, LEVELS4 = 3840, 2160, 6
METRIC_4K = "4K dense alignment frames/s (6-level image pyramid + all-pixel LM Solve per frame)"


def render_4k(seq_id, n_frames, device):
    """Keyframe (u8 image, exact inverse depth) and n_frames following frames (u8) of one 4K sequence; the
    camera advances 0.02 m per frame (a 100 Hz camera at KITTI speed) so that every frame aligns against the
    same keyframe.  Intrinsics are the reference's hard-coded KITTI constants (include/image_processing_global.h:
    35-36): they are compiled into its hot path, so a 4K input means this wide-angle camera for it too."""
    import torch
    from odometry_b200 import synth
    kw = dict(w=W4, h=H4, device=device, fx=synth.FX, cx=synth.CX, cy=synth.CY)
    sc = synth.make_scene(1000 + seq_id, 2)
    key, z = synth.render(sc, synth.pose_from_motion(), **kw)
    inv = (1.0 / z).to(torch.float32)
    frames = torch.empty((n_frames, H4, W4), dtype=torch.uint8, device=device)
    for f in range(n_frames):
        t = f + 1
        T = synth.pose_from_motion(0.0, 0.0, 0.02 * t, 0.0, math.radians(0.01 * t), 0.0)
        frames[f], _ = synth.render(sc, T, **kw)
    return key, inv, frames, (synth.FX, synth.CX, synth.CY)


def lm_params_4k(mod, K):
    """LM parameters of config 5 for `mod` (odometry_b200.capi or oracle): iters {10,20,30,30,30,30}, Huber 28."""
    if hasattr(mod, "default_lm_params"):
        p = mod.default_lm_params()
    else:
        p = mod.lm_params(sum_mode=mod.SUM_FLOAT_SEQ)
    for l, it in enumerate((10, 20, 30, 30, 30, 30)):
        p.max_iters[l] = it
    p.K.fx, p.K.cx, p.K.cy = K
    return p


def config_4k(args):
    return {"workload": "config 5: 3840x2160, 6-level image pyramid of every new frame + all-pixel dense LM alignment "
                        "against a keyframe with exact dense inverse depth (one sequence per GPU)",
            "frame": f"{W4}x{H4}", "levels": LEVELS4, "lm": "lambda 0.01, precision 0.995, iters {10,20,30,30,30,30}, Huber 28",
            "trajectory": "0.02 m/frame forward, yaw 0.01 deg/frame; every frame aligns against frame 0; warm start from the previous pose",
            "parallelism": f"one independent sequence per GPU x{args.gpus} GPUs, no collective",
            "l2": "one linearisation at level 0 streams 396 MB of point list (> 126 MB L2); no explicit flush"}


def run_reference_4k(args, rank):
    """--impl reference --workload 4k: the reference's Solve on 4K frames, one sequence per host thread."""
    if rank != 0:
        return
    import torch
    import oracle
    from oracle import ref
    cores = os.cpu_count() or 1
    n_seq = max(1, min(cores, 8))      # ~1.3 GB of Jacobian / weights per solve
    K_steps = max(1, min(args.steps, 1))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    use_ref = ref.available()
    data = []
    for s in range(n_seq):
        key, inv, frames, K = render_4k(s, 1, dev)
        data.append((key.cpu().numpy().astype(np.float32), inv.cpu().numpy(), frames[0].cpu().numpy().astype(np.float32)))
    p = lm_params_4k(oracle, K)
    out = [None] * n_seq

    def work(s):
        k, d, f = data[s]
        if use_ref:
            out[s] = ref.lm_solve(k, d, f, LEVELS4, p)
        else:
            o1, od, o2 = (oracle.image_pyramid(k, LEVELS4, True, packed=True), oracle.depth_pyramid(d, LEVELS4, False, packed=True),
                          oracle.image_pyramid(f, LEVELS4, True, packed=True))
            out[s] = oracle.lm_solve(o1, od, o2, W4, H4, LEVELS4, p)[:2]

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(s,)) for s in range(n_seq)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    fps = n_seq * K_steps / dt
    kind = "reference" if use_ref else "port"
    line = {"impl": "reference", "metric": METRIC_4K, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": K_steps, "warmup": 0, "ms_per_step": dt / K_steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_4k(args),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": n_seq, "kind": kind,
                             "sample": f"{n_seq} sequences x 1 frame (keyframe pyramids + frame pyramid + Solve), one per host thread; "
                                       + CPU_KIND_NOTE[kind]},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_4k(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    init_dist(dev, world)
    line = measure_4k(args, rank, local_rank, world, args.steps, args.warmup, not args.no_e2e, not args.no_cpu_baseline)
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def measure_4k(args, rank, local_rank, world, K, Wm, do_e2e, do_cpu):
    """config 5 on this rank's GPU; returns the bench line (also embedded in the default kitti line as config5_4k)"""
    import torch
    from odometry_b200 import capi, shard

    dev = torch.device("cuda", local_rank)
    key8, inv, frames8, Kcam = render_4k(rank, Wm + K, dev)
    from odometry_b200 import synth
    t_last = Wm + K
    T_gt_last = synth.relative_pose(synth.pose_from_motion(), synth.pose_from_motion(0.0, 0.0, 0.02 * t_last, 0.0,
                                                                                      math.radians(0.01 * t_last), 0.0))
    stream = torch.cuda.current_stream(dev)
    ctx = capi.Context(local_rank, stream.cuda_stream)   # the context runs on torch's stream: events below see its work
    p = lm_params_4k(capi, Kcam)
    key_f = key8.to(torch.float32)
    p1 = ctx.image_pyramid_dev(key_f.data_ptr(), W4, H4, LEVELS4, True)
    d1 = ctx.depth_pyramid_dev(inv.data_ptr(), W4, H4, LEVELS4, False)
    ctx.sync()
    frames_f = frames8.to(torch.float32)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    acc = {"pyr": 0.0, "lm": 0.0, "iters": np.zeros(LEVELS4, np.int64), "ok": True}

    def step_dev(f, T_prev, timed_stages):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record(stream)
        p2 = ctx.image_pyramid_dev(frames_f[f].data_ptr(), W4, H4, LEVELS4, True)
        e1.record(stream)
        rc, T, st = ctx.lm_solve(p1, d1, p2, p, T_prev)
        e2.record(stream)
        p2.release()
        if timed_stages:
            e2.synchronize()
            acc["pyr"] += e0.elapsed_time(e1)
            acc["lm"] += e1.elapsed_time(e2)
            acc["iters"] += np.array(list(st.iters)[:LEVELS4])
            acc["ok"] &= (rc == 0)
        return T

    def run(f0, f1, T, stepfn, timed_stages=False):
        for f in range(f0, f1):
            T = stepfn(f, T, timed_stages)
        return T

    def timed(fn):
        shard.barrier(dev)
        e0, e1 = ev(), ev()
        e0.record(stream)
        r = fn()
        e1.record(stream)
        shard.barrier(dev)
        return shard.max_over_ranks(e0.elapsed_time(e1), dev), r

    T0 = np.eye(4, dtype=np.float32)
    T = run(0, Wm, T0, step_dev)
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    sampler.start()
    ms_total, T_last = timed(lambda: run(Wm, Wm + K, T, step_dev, True))
    clocks = sampler.result()
    launches = ctx.launch_count() - launches0
    value = world * K / (ms_total * 1e-3)

    e2e = None
    if do_e2e:
        host = frames_f.cpu().pin_memory()

        def step_host(f, T_prev, _):
            p2 = ctx.image_pyramid(host[f].numpy(), LEVELS4, True)      # H2D of the 33 MB frame inside the call
            rc, Tn, st = ctx.lm_solve(p1, d1, p2, p, T_prev)           # D2H of the result inside the call
            p2.release()
            return Tn

        Tw = run(0, Wm, T0, step_host)
        ms_e2e, _ = timed(lambda: run(Wm, Wm + K, Tw, step_host))
        e2e = {"value": world * K / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": W4 * H4 * 4 + 64,
               "d2h_bytes_per_step": 1680, "ms_per_step": ms_e2e / K,
               "host_buffers": "pinned float32 (CV_32F) frame", "api": "odo_image_pyramid_build + odo_lm_solve"}
        del host

    hbm_peak, peak_src = load_peaks()
    lvl_px = [(W4 >> l) * (H4 >> l) for l in range(LEVELS4)]
    iters = acc["iters"] / K
    lm_bytes = float(sum(12 * lvl_px[l] * iters[l] for l in range(LEVELS4)))   # 12 B per pixel per linearisation
    pyr_bytes = 4 * (lvl_px[0] + sum(lvl_px))
    lm_ms, pyr_ms = acc["lm"] / K, acc["pyr"] / K
    ach = lm_bytes / (lm_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "k_lm_solve", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            "traffic": None, "peak_source": peak_src, "launch_ms": lm_ms, "algorithmic_bytes_per_launch": lm_bytes,
            "linearisations_per_level": [float(x) for x in iters],
            "list_bytes_per_launch": float(sum((48 + 8) * max(((W4 >> l) - 8) * ((H4 >> l) - 8), 0) * iters[l] for l in range(LEVELS4))),
            "stage_ms_per_step": {"image_pyramid": pyr_ms, "lm_solve": lm_ms},
            "per_stage_gbs": {"image_pyramid": pyr_bytes / (pyr_ms * 1e-3) / 1e9, "lm_solve": ach}}

    cpu = None
    if rank == 0 and world == 1 and do_cpu:
        import oracle
        from oracle import ref
        po = lm_params_4k(oracle, Kcam)
        k_np, d_np, f_np = key_f.cpu().numpy(), inv.cpu().numpy(), frames_f[Wm].cpu().numpy()
        t0 = time.perf_counter()
        if ref.available():
            ref.lm_solve(k_np, d_np, f_np, LEVELS4, po)
            kind = "reference"
        else:
            oracle.lm_solve(oracle.image_pyramid(k_np, LEVELS4, True, packed=True), oracle.depth_pyramid(d_np, LEVELS4, False, packed=True),
                            oracle.image_pyramid(f_np, LEVELS4, True, packed=True), W4, H4, LEVELS4, po)
            kind = "port"
        dt = time.perf_counter() - t0
        cpu = {"value": 1.0 / dt, "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": "1 frame (3 pyramid builds + one cold-started Solve), 1 thread; " + CPU_KIND_NOTE[kind]}

    line = {"metric": METRIC_4K, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_4k(args), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "frames_per_step": world, "all_status_ok": bool(acc["ok"]),
            "final_pose_translation": [float(x) for x in T_last[:3, 3]],
            "pose_error_vs_gt_m": float(np.linalg.norm(T_last[:3, 3] - T_gt_last[:3, 3]))}
    p1.release()
    d1.release()
    ctx.close()
    return line


# ---------------------------------------------------------------------------------------------------------------------
# --workload rectify: the camera front end.  A step = UndistortRectify (src/camera.cpp:72-83) of a batch
# of raw 640x480 u8 stereo pairs into CV_32F rectified pairs (the convertTo of run_odometry_kitti_offline.cpp:348 fused),
# with the reference's own calibration (calibration_file/camchain.yaml, values kept in tests/golden/camera_patterns.py)
# and the stereoRectify results of the committed golden fixture.
METRIC_RECT = "undistort+rectify stereo pairs/s (640x480 u8 raw -> CV_32F rectified, both eyes)"
RW, RH = 640, 480


def rectify_setup():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import camera_patterns as cp
    G = np.load(os.path.join(ROOT, "tests", "golden", "cv2_camera.npz"))
    return cp, G


def config_rect(args, pairs):
    return {"workload": f"front end (SURVEY 8 f3): UndistortRectify of {pairs} raw 640x480 u8 stereo pairs per step per GPU into "
                        "CV_32F, INTER_LINEAR / BORDER_CONSTANT, tables from ConfigureCamera on the reference's calibration file",
            "pairs_per_gpu": pairs, "frame": f"{RW}x{RH}", "parallelism": f"independent pairs x{args.gpus} GPUs, no collective",
            "l2": f"each step reads {pairs * 2 * RW * RH / 1e6:.0f} MB of fresh raw frames and writes {pairs * 8 * RW * RH / 1e6:.0f} MB (> 126 MB L2)"}


def raw_pairs(n, seed):
    rng = np.random.default_rng(20190117 + seed)
    base = rng.integers(0, 256, (RH // 8 + 2, RW // 8 + 2)).astype(np.float32)
    out = np.empty((n, 2, RH, RW), np.uint8)
    yy, xx = np.mgrid[0:RH, 0:RW]
    for i in range(n):
        for e in range(2):
            sh = (i * 7 + e * 3) % 8
            out[i, e] = (base[(yy + sh) // 8, (xx + sh) // 8] * 0.75 + ((xx * 3 + yy * 5 + i) % 64)).astype(np.uint8)
    return out


def cpu_rectify_pairs_per_s(raw, maps, threads):
    import oracle
    n = raw.shape[0]
    t0 = time.perf_counter()

    def work(i):
        for e in range(2):
            oracle.remap_bilinear(raw[i, e], maps[e][0], maps[e][1], 0).astype(np.float32)

    if threads == 1:
        for i in range(n):
            work(i)
    else:
        ths = [threading.Thread(target=lambda k=k: [work(i) for i in range(k, n, threads)]) for k in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    return n / (time.perf_counter() - t0)


def run_reference_rectify(args, rank):
    if rank != 0:
        return
    import oracle
    cp, G = rectify_setup()
    maps = [oracle.init_undistort_rectify_map(G[K], G[D], G[R], G[P], cp.RAW_SIZE) for K, D, R, P in
            (("K0", "D0", "R1", "P1"), ("K1", "D1", "R2", "P2"))]
    cores = os.cpu_count() or 1
    n = 16 * cores
    raw = raw_pairs(n, 0)
    K_steps = max(1, min(args.steps, 5))
    cpu_rectify_pairs_per_s(raw[:cores], maps, cores)
    t0 = time.perf_counter()
    for _ in range(K_steps):
        cpu_rectify_pairs_per_s(raw, maps, cores)
    dt = time.perf_counter() - t0
    v = n * K_steps / dt
    line = {"impl": "reference", "metric": METRIC_RECT, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": K_steps,
            "warmup": 1, "ms_per_step": dt / K_steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config_rect(args, n),
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{n} pairs per step, one thread per logical core; the oracle's restatement of cv::remap "
                                       "(the reference calls OpenCV here; oracle/camera.cpp is pinned bit-exactly to cv2 4.13.0); "
                                       "host: " + host_cpu()},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)


def run_rectify(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from odometry_b200 import capi, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    init_dist(dev, world)
    K, Wm = args.steps, args.warmup
    cp, G = rectify_setup()
    stream = torch.cuda.Stream(device=dev)   # a real (non-null) stream: the context runs on it and the events below see its work
    ctx = capi.Context(local_rank, stream.cuda_stream)
    cams = []
    for Kk, D, R, P in (("K0", "D0", "R1", "P1"), ("K1", "D1", "R2", "P2")):
        cam = capi.Camera(ctx, 4, G[Kk], G[D], sensor=(5.76, 4.29), resolution=(640, 482))
        cam.configure(G[R], G[P], cp.RAW_SIZE)
        cams.append(cam)
    B = 256
    nbuf = 3   # rotating input batches: every step reads raw frames that are not in L2
    raw_np = raw_pairs(B, rank)
    raws = [torch.from_numpy(np.roll(raw_np, k, axis=0).copy()).to(dev) for k in range(nbuf)]   # [B][2][H][W] u8
    out = torch.empty((2, B, RH, RW), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    npx = RW * RH
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(i):
        r = raws[i % nbuf]
        for e in range(2):
            cams[e].undistort_rectify_dev(r.data_ptr() + e * npx, True, RW, RH, 2 * npx, out[e].data_ptr(), False, npx, B)

    def timed(fn):
        shard.barrier(dev)
        e0, e1 = ev(), ev()
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        shard.barrier(dev)
        return shard.max_over_ranks(e0.elapsed_time(e1), dev)

    for i in range(Wm):
        step(i)
    sampler = ClockSampler(local_rank)
    l0 = ctx.launch_count()
    sampler.start()
    ms_total = timed(lambda: [step(Wm + i) for i in range(K)])
    clocks = sampler.result()
    launches = ctx.launch_count() - l0
    value = world * B * K / (ms_total * 1e-3)

    e2e = None
    if not args.no_e2e:
        ne = 64
        host = torch.from_numpy(raw_np[:ne]).pin_memory()
        host_out = torch.empty((2, ne, RH, RW), dtype=torch.float32).pin_memory()

        def step_host():
            for e in range(2):   # per eye: H2D of the raw frames + one launch + D2H of the rectified images inside the call
                cams[e].undistort_rectify_batch_ptrs([host[i, e].data_ptr() for i in range(ne)], True, RW, RH,
                                                     [host_out[e, i].data_ptr() for i in range(ne)], False)

        step_host()
        ms = timed(lambda: [step_host() for _ in range(K)])
        e2e = {"value": world * ne * K / (ms * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": ne * 2 * npx,
               "d2h_bytes_per_step": ne * 2 * npx * 4, "ms_per_step": ms / K, "pairs_per_step": ne,
               "host_buffers": "pinned u8 raw frames, pinned float32 results",
               "api": "odo_camera_undistort_rectify_batch (one call per eye; the rectified CV_32F images are the result read back: "
                      "PCIe-bound, 5 bytes per pixel)"}

    hbm_peak, peak_src = load_peaks()
    launch_ms = ms_total / K / 2
    alg = B * npx * (1 + 4) + 8 * npx      # raw u8 in, CV_32F out, the float2 table once per launch
    ach = alg / (launch_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "k_remap<u8>", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
            "traffic": None, "peak_source": peak_src, "launch_ms": launch_ms, "algorithmic_bytes_per_launch": alg,
            "images_per_launch": B}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle
        maps = [c.maps() for c in cams]
        n = 24
        cpu_rectify_pairs_per_s(raw_np[:2], maps, 1)
        v = cpu_rectify_pairs_per_s(raw_np[:n], maps, 1)
        cpu = {"value": v, "unit": "pairs/s", "cores": 1, "kind": "port",
               "sample": f"{n} pairs, 1 thread; the oracle's restatement of cv::remap (pinned bit-exactly to cv2 4.13.0); host: " + host_cpu()}
        ctx.sync()
        chk = out[0, 0].cpu().numpy()
        ref0 = oracle.remap_bilinear(raws[(Wm + K - 1) % nbuf][0, 0].cpu().numpy(), maps[0][0], maps[0][1], 0).astype(np.float32)
        roof["checked_against_oracle"] = bool(np.array_equal(chk, ref0))
    if rank == 0:
        line = {"metric": METRIC_RECT, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config_rect(args, B), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "pairs_per_step": world * B}
        emit(line)
    for c in cams:
        c.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    stdout_to_stderr()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if args.workload == "4k":
            run_reference_4k(args, rank)
        elif args.workload == "rectify":
            run_reference_rectify(args, rank)
        else:
            run_reference(args, rank, world)
        return
    if args.workload == "4k":
        run_4k(args, rank, local_rank, world)
        return
    if args.workload == "rectify":
        run_rectify(args, rank, local_rank, world)
        return

    run_kitti(args, rank, local_rank, world)


def run_kitti(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from odometry_b200 import capi, shard

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    placement = bind_to_gpu_numa(local_rank)     # before any pinned allocation and before NCCL spawns its threads
    init_dist(dev, world)

    (S, scaling), K, Wm = seqs_per_gpu(args, world), args.steps, args.warmup
    n_frames = 1 + Wm + K
    # config 4: world*S independent sequences, sequence s on rank s mod world (odometry_b200/shard.py)
    seq_ids = shard.shard_sequences(world * S, world, rank)
    frames_u8 = render_sequences(seq_ids, n_frames, dev, args.scene)    # [F][S][2][H][W] u8 on device
    gt = gt_relative_poses(n_frames)
    stream = torch.cuda.current_stream(dev)
    # T trackers per GPU, each with its own context (= its own CUDA streams) and S/T of the sequences.  A
    # frame is a dependent chain of kernels, some throughput bound (disparity search), some latency bound
    # (the LM and refinement iteration loops); independent trackers interleave on the GPU so that one
    # tracker's latency-bound stages run under another's throughput-bound ones.  Pure orchestration of the
    # same C-ABI calls: sequences are sharded over streams exactly as they are over GPUs.
    T = max(1, min(args.streams, S))
    while S % T:
        T -= 1
    St = S // T
    ctxs = [capi.Context(local_rank) for _ in range(T)]
    ctx = ctxs[0]

    def barrier():
        shard.barrier(dev)

    def max_over_ranks(ms):
        return shard.max_over_ranks(ms, dev)

    def launch_total():
        return sum(c.launch_count() for c in ctxs)

    def run_pipelined(trks, enqueue, f0, f1, on_result=None):
        """Two frames in flight per tracker: frame f+1 is enqueued (copies + kernels) before frame f is
        collected; every frame's poses are read back."""
        for i, t in enumerate(trks):
            enqueue(t, i, f0)
        for f in range(f0, f1):
            for i, t in enumerate(trks):
                if f + 1 < f1:
                    enqueue(t, i, f + 1)
                res = t.collect()
                assert res[0] == 0, "tracking failed"
                if on_result:
                    on_result(t, i, f, res)

    def timed(fn):
        """Device time of fn() between two idle points: e0 is recorded with the GPU idle, e1 after the host
        has collected the last result of every tracker (nothing is left running)."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    # ---------------- leg 1: inputs resident in HBM (CV_8U, the reference's input format) ----------------
    left_dev = frames_u8[:, :, 0].contiguous()                          # [F][S][H][W] u8 on device
    right_dev = frames_u8[:, :, 1].contiguous()
    trks = [capi.Tracker(c, W, H, LEVELS, St) for c in ctxs]

    def enq_dev(t, i, f):
        t.enqueue_dev(left_dev[f, i * St].data_ptr(), right_dev[f, i * St].data_ptr(), u8=True)

    run_pipelined(trks, enq_dev, 0, 1 + Wm)
    sampler = ClockSampler(local_rank)
    stage_acc, tally = {}, {"kf": 0, "ok": True}
    final_pose = np.zeros((S, 4, 4), np.float32)

    def on_result(t, i, f, res):
        rc, pa, pk, kf, st = res
        tally["kf"] += int(kf.sum())
        tally["ok"] &= bool((st == 0).all())
        if f == n_frames - 1:
            final_pose[i * St:(i + 1) * St] = pa
        for k, v in t.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / T

    launches0 = launch_total()
    sampler.start()
    ms_total = timed(lambda: run_pipelined(trks, enq_dev, 1 + Wm, n_frames, on_result))
    clocks = sampler.result()
    launches = launch_total() - launches0
    value = world * S * K / (ms_total * 1e-3)
    n_kf = int(shard.sum_over_ranks(tally["kf"], dev))
    status_ok = bool(shard.sum_over_ranks(0.0 if tally["ok"] else 1.0, dev) == 0.0)
    pose_err = pose_error_summary(final_pose, gt[n_frames - 1])
    pose_err["max"] = shard.max_over_ranks(pose_err["max"], dev)
    pose_err["max_rel"] = shard.max_over_ranks(pose_err["max_rel"], dev)
    pose_err["note"] = (f"translation error of pose_abs after frame {n_frames - 1} of every sequence (device-resident leg); median of "
                        "rank 0's sequences, max over all ranks; the reference itself drifts by about 1 % of the path (floor() "
                        "sampling), see pose_error_vs_gt of --impl reference")
    for t in trks:
        t.close()
    del left_dev, right_dev

    # ---------------- leg 2: end to end through the C-ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        def e2e_leg(host, u8):
            trks = [capi.Tracker(c, W, H, LEVELS, St) for c in ctxs]
            # one host pointer per image, as a C++ caller holds them (cv::Mat::data of every frame): the arrays are built
            # once, outside the timed region -- indexing a torch tensor 128 times per step costs more than the GPU step
            esz = host.element_size()
            p0 = host.data_ptr()
            arrs = [[(t.pointer_array([p0 + esz * (((f * S + i * St + s) * 2 + 0) * H * W) for s in range(St)]),
                      t.pointer_array([p0 + esz * (((f * S + i * St + s) * 2 + 1) * H * W) for s in range(St)]))
                     for i, t in enumerate(trks)] for f in range(n_frames)]

            def enq_host(t, i, f):
                t.enqueue_arrays(arrs[f][i][0], arrs[f][i][1], u8=u8)

            run_pipelined(trks, enq_host, 0, 1 + Wm)
            ms = timed(lambda: run_pipelined(trks, enq_host, 1 + Wm, n_frames))
            for t in trks:
                t.close()
            return ms

        # the frames as the reference's imread(GRAYSCALE) leaves them: u8, one pinned buffer per image
        host8 = frames_u8.cpu().pin_memory()                            # [F][S][2][H][W] u8
        ms_e2e = e2e_leg(host8, True)
        del host8
        h2d = S * 2 * W * H
        e2e = {"value": world * S * K / (ms_e2e * 1e-3), "unit": "frames/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": S * 240,
               "ms_per_step": ms_e2e / K, "host_buffers": "pinned u8 (what imread(GRAYSCALE) returns), one pointer per image",
               "api": "odo_tracker_enqueue_u8 + odo_tracker_collect (two frames in flight per tracker: the copy of "
                      "frame f+1 overlaps the kernels of frame f; the convertTo(CV_32F) of run_odometry_kitti_offline.cpp:348 "
                      "runs on the device; every frame's poses are read back)",
               "frac_of_device_rate": ms_total / ms_e2e,
               "h2d_gbs_achieved_per_rank": h2d * K / (ms_e2e * 1e-3) / 1e9,
               "h2d_gbs_probe_this_rank": h2d_probe_gbs(dev, stream),
               "host_placement": placement}
        host = frames_u8.to(torch.float32).cpu().pin_memory()           # the same call with CV_32F frames (4x the bytes)
        ms_f32 = e2e_leg(host, False)
        del host
        e2e["f32_ingest"] = {"value": world * S * K / (ms_f32 * 1e-3), "unit": "frames/s",
                             "h2d_bytes_per_step": S * 2 * W * H * 4, "api": "odo_tracker_enqueue (CV_32F host frames)"}

    # ---------------- roofline of the dominant kernel (device-resident leg, this rank) ----------------
    # Stage times are CUDA-event intervals on the stream each stage runs on, averaged over the timed steps.
    # The LM solve runs on its own stream CONCURRENTLY with blur_select + disparity (+ refine), so the stages
    # do not add up to the step; the critical path is image_pyramid -> blur_select -> disparity ->
    # depth_refine -> prepare, and the dominant kernel is taken among those.
    hbm_peak, peak_src = load_peaks()
    stage_ms = {k: v / K for k, v in stage_acc.items()}
    npx = W * H
    pyr_px = sum((W >> l) * (H >> l) for l in range(LEVELS))
    # algorithmic (compulsory) bytes per frame of each stage, DESIGN.md section 3
    alg_bytes = {
        "image_pyramid": npx + 4 * pyr_px,                   # u8 frame in, float levels out
        "blur_select": 5 * npx + 4 * npx + npx,              # blur R (u8 read, float write), gradient read of blur L, val
        "disparity": 17 * npx,                               # read L,R blur, val; write disp, dep
        "depth_refine": 20 * 40960,                          # one pass of the slot state (11-15 passes run)
        "depth_pyr_promote_prepare": 8 * pyr_px,             # point lists of a promoted keyframe: read img + depth
    }
    # (image_pyramid: the pyramid of frame k+1 is built on a third stream while frame k is still in flight -- its event
    # interval is a span on a low-priority stream, not a kernel time, and it is off the critical path)
    critical = [k for k in stage_ms if k not in ("lm_solve", "image_pyramid")]
    dom = max(critical, key=lambda k: stage_ms[k])
    traffic = None
    prof = {}
    try:
        import glob
        prof = json.load(open(sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*_traffic.json")), key=os.path.getmtime)[-1]))  # newest capture
        if prof.get("sequences_per_launch") == St and dom in prof.get("stage_dram_bytes", {}):
            traffic = prof["stage_dram_bytes"][dom]
    except Exception:
        pass
    # a launch of the dominant kernel covers the St sequences of one tracker
    ach = alg_bytes[dom] * St / (stage_ms[dom] * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": {"disparity": "k_disparity<true,1280,prune>"}.get(dom, dom), "stage": dom,
            "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
            "peak_source": peak_src, "launch_ms": stage_ms[dom],
            "algorithmic_bytes_per_launch": alg_bytes[dom] * St, "sequences_per_launch": St,
            "stage_ms_per_step": stage_ms,
            "stages_overlap": "lm_solve runs on a second stream concurrently with blur_select/disparity/depth_refine; "
                              "image_pyramid of the NEXT frame runs on a third stream under all of them (its entry is a span, not a kernel time)",
            "per_stage_gbs": {k: alg_bytes[k] * St / (stage_ms[k] * 1e-3) / 1e9 for k in alg_bytes if stage_ms.get(k) and k != "image_pyramid"},
            "hbm_note": "frac = algorithmic HBM bytes of the dominant kernel / its time / HBM peak (the contract's definition); "
                        "the line search is bound by instruction issue, not HBM: see issue_roofline"}
    # the kernels of the critical path are bound by instruction issue, not by HBM: achieved warp instructions per second
    # (instruction count of the committed ncu capture of the same launch shape / the live launch time) against the
    # issue slots of the device (SMs x 4 schedulers x clock)
    try:
        kname = {"disparity": "k_disparity<1, 1280, 1>", "depth_refine": "k_refine", "blur_select": "k_select_tile<28>"}.get(dom)
        kinfo = prof.get("kernels", {}).get(kname) if prof.get("sequences_per_launch") == St else None
        if kinfo and kinfo.get("warp_inst_per_launch"):
            sm = torch.cuda.get_device_properties(dev).multi_processor_count
            clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
            peak_issue = sm * 4 * clk
            # the stage may hold small helper kernels besides the dominant one: scale the stage time by the kernel's share
            share = kinfo["ms_per_launch"] / max(prof["stage_ms_serialised"].get(dom, kinfo["ms_per_launch"]), 1e-9)
            t_k = stage_ms[dom] * min(share, 1.0) * 1e-3
            roof["issue"] = {"bound": "issue", "kernel": kname, "unit": "warp instructions/s",
                             "warp_instructions_per_launch": kinfo["warp_inst_per_launch"],
                             "achieved": kinfo["warp_inst_per_launch"] / t_k, "peak": peak_issue,
                             "frac": kinfo["warp_inst_per_launch"] / t_k / peak_issue,
                             "ncu_issue_active_pct": kinfo.get("issue_active_pct"),
                             "source": prof.get("source")}
    except Exception as e:
        roof["issue"] = {"error": str(e)}
    if dom == "disparity":
        # the line search is bound by integer-multiply issue, not by HBM: 8 IMAD per (pixel, candidate) on the
        # half-rate fmaheavy pipe (ncu: sm__pipe_fmaheavy_cycles_active).  Candidate evaluations of one frame are
        # counted from the selection mask of a frame of the timed region (through the C-ABI).
        try:
            f_probe = 1 + Wm
            rc_, _, _, val_ = ctx.select_pixels(frames_u8[f_probe, 0, 0].float().cpu().numpy(),
                                                frames_u8[f_probe, 0, 1].float().cpu().numpy(), capi.default_depth_params())
            ys, xs = np.nonzero(val_)
            evals = float(np.maximum(xs - 4, 0).sum())
            sm = torch.cuda.get_device_properties(dev).multi_processor_count
            clk = (clocks.get("sm_mhz") or 1965.0) * 1e6
            peak_evals = sm * 4 * 16 * clk / 8.0              # 16 IMAD lanes per scheduler per clock, 8 IMAD per evaluation
            ach_evals = evals * St / (stage_ms[dom] * 1e-3)
            roof["issue_roofline"] = {"bound": "issue", "pipe": "imad (fmaheavy)", "unit": "candidate evaluations/s",
                                      "note": "ALGORITHMIC-work figure: evaluations = the reference's (pixel, candidate) count, "
                                              "peak = 8 IMAD each at the fmaheavy issue rate; the pruned scan does NOT perform "
                                              "most of them (it rejects whole chunks from a 2-IDP4A bound), so this is achieved "
                                              "algorithmic work per second against the brute-force issue peak, not pipe utilisation "
                                              "(ncu: profiles/ncu_*_summary.md)",
                                      "evaluations_per_frame": evals, "achieved": ach_evals, "peak": peak_evals,
                                      "frac": ach_evals / peak_evals}
        except Exception as e:  # the probe is informational
            roof["issue_roofline"] = {"error": str(e)}

    # ---------------- extras: single-sequence latency (north-star target < 1 ms) and config 5 ----------------
    latency = None
    cfg5 = None
    if not args.no_extras:
        trk1 = capi.Tracker(ctx, W, H, LEVELS, 1)
        l1 = frames_u8[:, 0, 0].contiguous()
        r1 = frames_u8[:, 0, 1].contiguous()
        for f in range(0, 1 + Wm):
            assert trk1.step_dev(l1[f].data_ptr(), r1[f].data_ptr(), u8=True)[0] == 0

        def lat():
            for f in range(1 + Wm, n_frames):
                assert trk1.step_dev(l1[f].data_ptr(), r1[f].data_ptr(), u8=True)[0] == 0     # one frame at a time: enqueue + collect

        ms_lat = timed(lat)
        latency = {"latency_s1_ms": ms_lat / K, "frames": K,
                   "what": "one 1241x376 stereo frame of ONE sequence, step by step (odo_tracker_step_dev_u8: the pose is read back "
                           "before the next frame is queued): pyramid + Solve + ComputeDepth + keyframe rule"}
        trk1.close()
        del l1, r1
    del frames_u8
    if not args.no_extras:
        torch.cuda.empty_cache()
        l5 = measure_4k(args, rank, local_rank, world, 5, 3, False, False)
        cfg5 = {"value": l5["value"], "unit": l5["unit"], "ms_per_step": l5["ms_per_step"], "steps": 5, "warmup": 3,
                "roofline": l5["roofline"], "all_status_ok": l5["all_status_ok"], "pose_error_vs_gt_m": l5["pose_error_vs_gt_m"],
                "config": l5["config"]["workload"]}

    # ---------------- CPU baseline (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ns = min(args.cpu_seqs, S)
        nf = min(n_frames, 1 + 30)
        fr = render_sequences(seq_ids[:ns], nf, dev, args.scene).cpu().numpy()
        fps, nfr, kind = cpu_tracking_fps(fr, ns, nf, threads=1)
        cpu = {"value": fps, "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": f"{ns} sequences x {nf - 1} frames of the same workload, 1 thread "
                         f"(host has {os.cpu_count()} logical cores); " + CPU_KIND_NOTE[kind]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "frames_per_step": world * S, "keyframes_in_timed_region": n_kf,
            "keyframe_fraction": n_kf / float(world * S * K), "all_status_ok": status_ok,
            "pose_error_vs_gt": pose_err,
            "disparity_mpix_per_s": (npx * S / 1e6) / ((stage_ms["blur_select"] + stage_ms["disparity"] + stage_ms["depth_refine"]) * 1e-3),
        }
        if latency:
            line.update(latency)
        if cfg5:
            line["config5_4k"] = cfg5
        emit(line)
    for c in ctxs:
        c.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
