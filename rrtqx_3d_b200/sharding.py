"""Sharding of independent query / edge batches over ranks (one process per GPU).

The tree and the obstacle set are replicated; rank g owns the contiguous slice
[g*n/G, (g+1)*n/G) of a batch.  The only collective is the gather of fixed-size
results (per-query counts, nearest idx/dist, per-edge flags); neighbour lists stay
sharded.  Works with any torch.distributed backend: NCCL on
the GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous slice of rank `rank`: [lo, hi)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_sizes(n: int, world: int):
    return [shard_bounds(n, g, world)[1] - shard_bounds(n, g, world)[0] for g in range(world)]


def gather_fixed(local, n_total: int, dist=None, device=None):
    """All-gather a per-item result (1-D array / tensor of this rank's slice) into the full
    array of n_total items on every rank.  `dist` = torch.distributed (initialised)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=t.dtype, device=t.device)
    buf[: t.numel()] = t
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    full = torch.cat([o[:s] for o, s in zip(out, sizes)])
    return full if isinstance(local, torch.Tensor) else full.cpu().numpy()
