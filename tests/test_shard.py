"""Multi-GPU path = sharding of independent sequences, no data-path collective.  The host-side logic
(partition, barrier, max-over-ranks timing, pose gather) is covered here with world_size-2 gloo
processes on CPU; the GPU test checks that a sequence gives bit-identical poses in any shard."""
import os
import socket

import numpy as np
import pytest

from odometry_b200 import shard


def test_partition_is_exact_and_balanced():
    for n_total in (1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            parts = [shard.shard_sequences(n_total, world, r) for r in range(world)]
            flat = sorted(s for p in parts for s in p)
            assert flat == list(range(n_total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            for r, p in enumerate(parts):
                assert all(shard.owner_of(s, world) == r for s in p)
    with pytest.raises(ValueError):
        shard.shard_sequences(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ids = shard.shard_sequences(n_total, world, rank)
        # a fake "pose" that encodes the global sequence id, as each rank's tracker would return
        local = np.stack([np.full((4, 4), float(s), np.float32) for s in ids]) if ids else np.zeros((0, 4, 4), np.float32)
        shard.barrier()
        slowest = shard.max_over_ranks(10.0 + rank)
        total = shard.sum_over_ranks(len(ids))
        allp = shard.gather_poses(local, n_total)
        q.put((rank, ids, slowest, total, None if allp is None else allp[:, 0, 0].tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 8])
def test_world_size_2_gloo(n_total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ids0, slow0, tot0, all0), (r1, ids1, slow1, tot1, all1) = res
    assert sorted(ids0 + ids1) == list(range(n_total))
    assert slow0 == slow1 == 11.0            # the job takes as long as its slowest shard
    assert tot0 == tot1 == n_total
    assert all0 == [float(s) for s in range(n_total)] and all1 is None


@pytest.mark.gpu
def test_sequence_result_does_not_depend_on_its_shard(ctx, synth_data):
    """Sequence 1 tracked inside the batch of 'rank 1 of 2' ({1, 3}) equals the same sequence tracked
    inside the single-GPU batch ({0, 1, 2, 3}) bit for bit: sharding changes nothing but the speed."""
    from odometry_b200 import capi
    seqs = [synth_data.sequence(s, n_frames=3) for s in range(4)]

    def run(ids):
        t = capi.Tracker(ctx, 1241, 376, 4, len(ids))
        out = []
        for k in range(3):
            rc, pa, pk, kf, st = t.step([seqs[s]["frames"][k][0] for s in ids], [seqs[s]["frames"][k][1] for s in ids])
            assert rc == 0
            out.append((pa, pk, kf))
        t.close()
        return out

    full = run(shard.shard_sequences(4, 1, 0))
    for rank in range(2):
        ids = shard.shard_sequences(4, 2, rank)
        part = run(ids)
        for k in range(3):
            for j, s in enumerate(ids):
                assert np.array_equal(part[k][0][j], full[k][0][s])
                assert np.array_equal(part[k][1][j], full[k][1][s])
                assert part[k][2][j] == full[k][2][s]
