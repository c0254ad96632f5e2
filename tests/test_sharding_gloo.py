"""The N > 1 path on CPU: world_size-2 gloo processes exercise shard assignment and the
max-over-ranks / whole-job throughput reductions that bench.py uses with nccl."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audio_transformers_b200 import sharding


def test_shard_range_tiles_everything():
    for n in (0, 1, 7, 64, 1000, 65536):
        for ws in (1, 2, 3, 4, 8):
            spans = [sharding.shard_range(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        assert sharding.world() == (rank, world_size)
        lo, hi = sharding.shard_range(n_items, rank, world_size)
        # each rank "processes" its shard: here a checksum of the clip indices, and a fake device time
        local_sum = float(sum(range(lo, hi)))
        fake_ms = 10.0 * (rank + 1)
        dist.barrier()
        total = sharding.sum_over_ranks(local_sum)
        worst = sharding.max_over_ranks(fake_ms)
        thr = sharding.job_throughput(hi - lo, fake_ms)
        q.put((rank, lo, hi, total, worst, thr))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port, n_items, ws = _free_port(), 129, 2
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_items, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(ws))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 65), (65, 129)]
    for r in res:
        assert r[3] == float(sum(range(n_items)))           # all clips covered exactly once
        assert r[4] == 20.0                                  # slowest rank
        assert abs(r[5] - n_items / 0.020) < 1e-6            # all units / max time
