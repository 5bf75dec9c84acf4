"""Instance sharding across the GPUs of one box.

Every QP / NLP instance is independent, so the batch is cut into contiguous blocks of ceil(N/G) instances, one
process per GPU, each with its own handle; the sparsity pattern and the options are replicated.  There is no
collective on the hot path.  The only communication is one final gather of per-instance results
(status, iterations, objective, KKT residuals) through torch.distributed (NCCL on the GPUs, gloo in CPU tests).
"""
import numpy as np


def shard_range(n_instances: int, rank: int, world: int):
    """[begin, end) of the contiguous block owned by `rank`: ceil(N/G) instances per rank, the tail may be short."""
    per = -(-n_instances // world)
    b = min(n_instances, rank * per)
    return b, min(n_instances, b + per)


def shard_sizes(n_instances: int, world: int):
    return [shard_range(n_instances, r, world)[1] - shard_range(n_instances, r, world)[0] for r in range(world)]


def gather_results(local: dict, n_instances: int, group=None, device=None):
    """Gather per-instance result arrays of all ranks; returns the full arrays on every rank (all_gather).

    `local` maps names to numpy arrays whose first dimension is this rank's shard size.  Ragged tails are padded to
    ceil(N/G) for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {k: np.asarray(v) for k, v in local.items()}
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    per = -(-n_instances // world)
    out = {}
    for k, v in local.items():
        v = np.ascontiguousarray(v)
        pad = np.zeros((per,) + v.shape[1:], dtype=v.dtype)
        pad[: v.shape[0]] = v
        t = torch.from_numpy(pad)
        if device is not None:
            t = t.to(device)
        bufs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(bufs, t, group=group)
        full = torch.cat(bufs, 0).cpu().numpy()
        sizes = shard_sizes(n_instances, world)
        parts = [full[r * per: r * per + sizes[r]] for r in range(world)]
        out[k] = np.concatenate(parts, 0)
    return out
