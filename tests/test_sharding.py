"""Host-side multi-GPU logic on CPU: contiguous instance shards and the final result gather over a
world_size-2 gloo group (the hot path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from restartsqp_b200.sharding import gather_results, shard_range, shard_sizes


def test_shard_ranges_cover_all_instances():
    for n in (0, 1, 7, 8, 1000, 10 ** 6):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert sum(shard_sizes(n, world)) == n
            assert max(shard_sizes(n, world)) <= -(-n // world) if n else True


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b, e = shard_range(n, rank, world)
    idx = np.arange(b, e)
    local = dict(status=(20 + idx % 3).astype(np.int32), obj=np.sin(idx.astype(np.float64)),
                 kkt=np.stack([idx * 1.0, idx * 2.0], 1))
    full = gather_results(local, n)
    if rank == 0:
        q.put({k: v for k, v in full.items()})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [11, 64])
def test_gather_over_gloo_world2(n):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    idx = np.arange(n)
    assert (full["status"] == 20 + idx % 3).all()
    assert (full["obj"] == np.sin(idx.astype(np.float64))).all()
    assert full["kkt"].shape == (n, 2) and (full["kkt"][:, 1] == 2.0 * idx).all()
