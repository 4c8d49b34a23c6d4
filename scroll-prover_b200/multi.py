"""Host-side planning helpers for the one-process-per-GPU launch, and the CPU twin of the sharded MSM.

The PRODUCT's multi-GPU data path lives inside the C ABI (csrc/comm.cu: the context-owned NCCL communicator,
b200zk_msm_g1_sharded, b200zk_graph_evaluate_rows + b200zk_allgather_rows; Python: Context.comm_init, Srs.msm_sharded).  This module
holds what needs no device:
  * shard_range  — the contiguous point / row range of a rank; b200zk_shard_range must return the same partition
                   (tests/test_multi_rank.py holds the two together);
  * assign_jobs  — longest-processing-time placement of a proof layer's independent column jobs over ranks;
  * msm_sharded  — the same gather-and-sum flow as b200zk_msm_g1_sharded over ANY torch.distributed backend, with the local MSM
                   and the final sum passed in: the world-2 gloo tests run it on CPU with the oracle standing in for the device.
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
