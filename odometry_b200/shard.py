"""Sharding of independent stereo sequences over the GPUs of one box.

A frame pair is sequential in pose (run_odometry_kitti_offline.cpp:215-218,261-268), so a sequence
never leaves its GPU and the data path has NO collective: sequence s runs on rank s mod G, every
rank batches its own sequences into the same launches.  torch.distributed (nccl on the GPUs, gloo in
the CPU tests) is used only for the plumbing around the timed region: the barrier, the max-over-ranks
of the device time and the final gather of the 64-byte poses.
"""
from __future__ import annotations

import numpy as np


def shard_sequences(n_total: int, world: int, rank: int) -> list[int]:
    """Global ids of the sequences owned by `rank` (round robin: s -> s mod world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    return list(range(rank, n_total, world))


def owner_of(seq_id: int, world: int) -> int:
    return seq_id % world


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def barrier(device=None):
    """Device sync + process barrier + device sync (bench.py brackets the timed region with it)."""
    import torch
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)
    d = _dist()
    if d is not None and d.get_world_size() > 1:
        d.barrier()
    if device is not None and torch.cuda.is_available():
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device="cpu") -> float:
    """Slowest rank's device time: the job is done when the last shard is."""
    import torch
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device="cpu") -> float:
    import torch
    d = _dist()
    if d is None or d.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.SUM)
    return float(t.item())


def gather_poses(local_poses: np.ndarray, n_total: int, device="cpu") -> np.ndarray | None:
    """local_poses: [n_local][...] for this rank's shard_sequences(); returns [n_total][...] in global
    sequence order on rank 0 (None elsewhere).  Outside the timed region: 64 B per sequence and frame."""
    import torch
    d = _dist()
    local_poses = np.ascontiguousarray(local_poses, dtype=np.float32)
    tail = local_poses.shape[1:]
    if d is None or d.get_world_size() == 1:
        assert local_poses.shape[0] == n_total
        return local_poses
    world, rank = d.get_world_size(), d.get_rank()
    n_max = (n_total + world - 1) // world
    pad = np.zeros((n_max,) + tail, np.float32)
    pad[: local_poses.shape[0]] = local_poses
    mine = torch.from_numpy(pad).to(device)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    d.all_gather(bufs, mine)
    if rank != 0:
        return None
    out = np.zeros((n_total,) + tail, np.float32)
    for r in range(world):
        ids = shard_sequences(n_total, world, r)
        out[ids] = bufs[r].cpu().numpy()[: len(ids)]
    return out
