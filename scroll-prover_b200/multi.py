"""Multi-GPU host logic: one process per GPU (torch.distributed, NCCL on the GPU box / gloo in CPU tests).

Two strategies, both without a data-path collective on field data:
  * job fan-out  — independent columns / commits of one proof are distributed over ranks (assign_jobs);
  * point-range MSM sharding — rank r holds bases[lo_r:hi_r] resident and receives the matching scalar slice;
    each rank computes a full local Pippenger, the 96-byte partial points are all-gathered (NCCL has no G1
    reduction) and every rank adds them with b200zk_g1_sum.
"""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous point range of `rank` (the first n % world ranks get one extra point)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assign_jobs(costs, world: int):
    """Longest-processing-time greedy: returns per-rank lists of job indices."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda x: (load[x], x))
        out[r].append(i)
        load[r] += costs[i]
    return out


def msm_sharded(local_msm, g1_sum, scalars_shard, dist=None, device=None) -> np.ndarray:
    """local_msm(scalars) -> (12,) uint64 Jacobian partial of this rank's point range;
    g1_sum((world,12)) -> (12,) total.  `dist` is torch.distributed (already initialised) or None."""
    part = np.ascontiguousarray(local_msm(scalars_shard), dtype=np.uint64)
    if dist is None or dist.get_world_size() == 1:
        return part
    import torch

    world = dist.get_world_size()
    mine = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        mine = mine.to(device)
    gathered = torch.empty((world, 12), dtype=torch.int64, device=mine.device)
    dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1))
    return g1_sum(gathered.cpu().numpy().view(np.uint64))
