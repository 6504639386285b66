"""Multi-GPU plumbing of the bulk path: block partition of the particle ids and the one collective a
bulk run needs (sum of the per-step observable series).  torch.distributed only; works with nccl (GPU)
and gloo (CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int):
    """[first, last) global particle ids of a rank: contiguous blocks, sizes differ by at most one."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    base, rem = divmod(int(n_total), world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def allreduce_observables(series, group=None):
    """Sum the [steps][valleys][3] = {sum E, sum v.E_dir, count} partial series of all ranks in place
    (torch tensor; one collective per run or per chunk -- bulk runs need no per-step communication)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(series, op=dist.ReduceOp.SUM, group=group)
    return series


def finalize_observables(series, n_total=None):
    """{sum E, sum v.E, count} -> per-valley <E>, <v.E_dir>, occupation, as the reference's getAvgEnergy /
    getAvgDriftVelocity / getValleyOccupationProbability return them (valleys without particles -> 0)."""
    s = np.asarray(series, dtype=np.float64)
    cnt = s[..., 2]
    total = cnt.sum(axis=-1, keepdims=True) if n_total is None else float(n_total)
    with np.errstate(invalid="ignore", divide="ignore"):
        e = np.where(cnt > 0, s[..., 0] / cnt, 0.0)
        v = np.where(cnt > 0, s[..., 1] / cnt, 0.0)
        occ = cnt / total
    return e, v, occ
