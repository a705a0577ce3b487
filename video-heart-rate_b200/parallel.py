"""Multi-GPU plumbing: clips and windows are independent, so the path shards with no
data-path collective.  One process per GPU (torchrun); unit ``i`` goes
to rank ``i mod world``; the only exchange is one final gather of the BPM results (a few KB)
over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_units(n_units: int, rank: int, world: int):
    """Round-robin ownership: equal-sized units (config c4: 64 identical-size clips)."""
    return list(range(rank, n_units, world))


def shard_by_cost(costs, world: int):
    """Unequal units (config c5: windows of different resolution x length): sort by cost,
    deal greedily to the least-loaded rank.  -> list of index lists, one per rank."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(i)
        load[r] += costs[i]
    return [sorted(o) for o in out]


def gather_results(local: dict, n_units: int, width: int, device=None):
    """Final gather: ``local`` maps unit id -> float64 array of ``width`` values.  Rank 0
    returns the dense (n_units, width) array, other ranks None.  One all_gather of
    n_units*width doubles per rank (NaN = not mine), merged on rank 0."""
    import torch
    import torch.distributed as dist
    buf = np.full((n_units, width), np.nan)
    for u, v in local.items():
        buf[u] = np.asarray(v, dtype=np.float64).reshape(width)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return buf
    t = torch.as_tensor(buf)
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    if dist.get_rank() != 0:
        return None
    out = np.full((n_units, width), np.nan)
    for p in parts:
        a = p.cpu().numpy()
        m = ~np.isnan(a)
        out[m] = a[m]
    return out
