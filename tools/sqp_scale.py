"""BASELINE.json configs[4] on the GPUs of one box: N HS-scale instances (HS071, perturbed starts) sharded contiguously by
instance index across the ranks (restartsqp_b200.sharding), full SQP per instance, no collective on the path; one final gather.
  python tools/sqp_scale.py [N]            (single GPU)
  torchrun --nproc-per-node G tools/sqp_scale.py [N]"""
import sys, os, time, numpy as np
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import restartsqp_b200 as r
from restartsqp_b200 import sharding
from restartsqp_b200.nl_reader import AmplNLP, DeviceNLP
from restartsqp_b200.sqp_driver import BatchedSQP
from restartsqp_b200.sqp_device import DeviceBatchedSQP
Loop = BatchedSQP if os.environ.get('SQP_HOST_LOOP') else DeviceBatchedSQP
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
host = AmplNLP(os.path.join(R, "tests", "golden", "hs_nl", "hs071.nl")); dev = DeviceNLP(host, device=local)
lo, hi = sharding.shard_range(N, rank, world)
x0, _ = host.Get_starting_point(); xl, xu, _, _ = host.Get_bounds_info()
rng = np.random.default_rng(71000)
X = np.clip(x0 * (1 + 0.1 * rng.standard_normal((N, host.n))) + 0.1 * rng.standard_normal((N, host.n)), xl, xu)[lo:hi]
Loop(dev, x0=X[:256], device=local).Optimize()
if dist is not None: dist.barrier()
t0 = time.perf_counter()
res = Loop(dev, x0=X, device=local).Optimize()
dt = time.perf_counter() - t0
if dist is not None:
    import torch
    t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
    nopt = torch.tensor([int((res.exitflag == 0).sum())], device="cuda"); dist.all_reduce(nopt); nopt = int(nopt.item())
else:
    nopt = int((res.exitflag == 0).sum())
if rank == 0:
    print({"instances": N, "gpus": world, "seconds": dt, "sqp_solves_per_s": N / dt, "optimal": nopt, "per_rank": hi - lo, "loop": Loop.__name__}, flush=True)
if dist is not None: dist.destroy_process_group()
