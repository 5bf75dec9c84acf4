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


def _sqp_worker(rank, world, port, n, q):
    """One rank of BASELINE.json configs[4] on CPU: its contiguous shard of the perturbed HS071 starts through the SQP loop (the
    C oracle stands in for the device loop: one independent solve per instance, like the GPU path), then the final gather."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import restartsqp_b200 as r
    from restartsqp_b200.nl_reader import AmplNLP
    from oracle import oracle_py as orc
    from test_hs_suite import HS_DIR, perturbed_starts
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    host = AmplNLP(os.path.join(HS_DIR, "hs071.nl"))
    b, e = shard_range(n, rank, world)
    X = perturbed_starts(host, n, 0)[b:e]
    res = orc.SqpOracle(host, r.Options()).solve_batch(X, nthreads=1)
    full = gather_results(dict(x=res["x"], obj=res["obj"], exitflag=res["exitflag"], iters=res["iters"]), n)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sqp_solves_equal_the_unsharded_run():
    """Instances are independent: N starts cut into two contiguous shards, solved by two ranks and gathered, give exactly the
    arrays one process computes for all N (ragged tail included: N = 37)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import restartsqp_b200 as r
    from restartsqp_b200.nl_reader import AmplNLP
    from oracle import oracle_py as orc
    from test_hs_suite import HS_DIR, perturbed_starts
    n = 37
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sqp_worker, args=(k, 2, port, n, q)) for k in range(2)]
    for p in procs:
        p.start()
    full = q.get(timeout=180)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    host = AmplNLP(os.path.join(HS_DIR, "hs071.nl"))
    ref = orc.SqpOracle(host, r.Options()).solve_batch(perturbed_starts(host, n, 0), nthreads=1)
    assert (full["exitflag"] == ref["exitflag"]).all() and (full["iters"] == ref["iters"]).all()
    assert np.array_equal(full["x"], ref["x"]) and np.array_equal(full["obj"], ref["obj"])
    assert (full["exitflag"] == 0).all()
