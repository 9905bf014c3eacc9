"""Multi-GPU plumbing for the stepper: environments are independent units, so the
job shards them across ranks with NO data-path collective (the reference does the
same with `jax.pmap` over an env axis, RSR/train.py:232-235).  Only scalars that
describe the run (timings, sweep losses) are reduced / gathered."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import prng


def shard_keys(seed: int, envs_per_rank: int, rank: int, world: int) -> np.ndarray:
    """Reset keys of this rank: rows [rank*n, (rank+1)*n) of split(PRNGKey(seed), n*world).
    The union over ranks equals the single-process key set; shards are disjoint."""
    keys = prng.split(prng.PRNGKey(seed), envs_per_rank * world)
    return keys[rank * envs_per_rank:(rank + 1) * envs_per_rank]


def env_range(envs_per_rank: int, rank: int) -> range:
    return range(rank * envs_per_rank, (rank + 1) * envs_per_rank)


def reduce_max(values, device=None) -> list:
    """max over ranks of a few host scalars (device timings): the job time is the slowest rank's"""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_concat(local: torch.Tensor) -> torch.Tensor:
    """concatenate per-rank 1-D results (e.g. friction-sweep losses of a parameter shard)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    out = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(out, local.contiguous())
    return torch.cat(out)
