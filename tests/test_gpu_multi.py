"""Point-range sharded MSM over 2 real GPUs (one process per GPU, NCCL all-gather of the 96-byte partials,
b200zk_g1_sum).  Skipped unless at least 2 CUDA devices are visible (run with `gpurun --gpus 2`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import importlib

    import torch
    import torch.distributed as dist

    from oracle import oracle as O

    zk = importlib.import_module("scroll-prover_b200")
    multi = importlib.import_module("scroll-prover_b200.multi")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = zk.Context(rank)
    bases = O.fill_points_chain(n, 41, 4)
    scal = O.fill_fr(n, 42, witness_like=True)
    lo, hi = multi.shard_range(n, rank, world)
    srs = ctx.srs_register(bases[lo:hi])  # this rank's resident point range
    total = multi.msm_sharded(lambda s: srs.msm(s), ctx.g1_sum, scal[lo:hi], dist, device=torch.device("cuda", rank))
    exp = O.best_multiexp(scal, bases, threads=4)
    ok = bool(np.array_equal(O.g1_to_affine(total), O.g1_to_affine(exp)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()
    ctx.close()


def test_point_range_sharded_msm_two_gpus():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [mpctx.Process(target=_worker, args=(r, 2, port, (1 << 17) + 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(0, True), (1, True)]
