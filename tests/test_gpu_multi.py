"""Point-range sharded MSM INSIDE the C ABI over 2 / 4 / 8 real GPUs (BASELINE configs[3]):
one process per GPU, every context joins the context-owned NCCL communicator (b200zk_ctx_comm_init; the 128-byte
unique id travels over torch.distributed, the bootstrap channel of this test), every rank passes ONLY its slice of
the scalars to b200zk_msm_g1_sharded and must receive the oracle's point, bit for bit, on every rank.
Each case is skipped unless that many CUDA devices are visible (`gpurun --gpus N`)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

CASES = [((1 << 17) + 3, True), (1 << 16, False), (5, False), (0, False), (1000, True)]  # (n, witness_like)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import importlib

    import torch
    import torch.distributed as dist

    from oracle import oracle as O

    zk = importlib.import_module("scroll-prover_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ctx = zk.Context(rank)
    ctx.comm_init_torch(dist)
    assert ctx.comm_info() == (rank, world)
    ok = True
    nmax = max(n for n, _ in CASES)
    bases = O.fill_points_chain(nmax, 41, 4)
    srs = ctx.srs_register(bases)  # replicated SRS (>= 2^16 points: precomputed tables, sliced in place)
    for n, wl in CASES:
        scal = O.fill_fr(n, 42 + n, witness_like=wl)
        first, cnt = zk.shard_range(n, rank, world)
        got = srs.msm_sharded(scal[first:first + cnt], n)
        exp = O.best_multiexp(scal, bases[:n], threads=4)
        ok &= bool(np.array_equal(O.g1_to_affine(got), O.g1_to_affine(exp)))
        # the range building block alone, against the oracle on the same slice
        part = srs.msm_range(scal[first:first + cnt], first)
        ok &= bool(np.array_equal(O.g1_to_affine(part), O.g1_to_affine(O.best_multiexp(scal[first:first + cnt], bases[first:first + cnt], threads=2))))
        # every rank holds the same bytes
        t = torch.from_numpy(got.view(np.int64).copy()).cuda()
        allg = torch.empty((world, 12), dtype=torch.int64, device="cuda")
        dist.all_gather_into_tensor(allg.view(-1), t)
        ok &= bool((allg == allg[0]).all().item())
    q.put((rank, ok))
    dist.barrier()
    srs.release()
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_point_range_sharded_msm_in_the_abi(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp

    mpctx = mp.get_context("spawn")
    q = mpctx.Queue()
    port = 29700 + (os.getpid() + 7 * world) % 1000
    procs = [mpctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=900) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(r, True) for r in range(world)]


def test_sharded_entry_points_degenerate_on_one_rank(ctx, zk):
    """world == 1: no communicator; b200zk_msm_g1_sharded == b200zk_msm_g1, ranges cover the SRS."""
    from oracle import oracle as O

    n = 3000
    bases = O.fill_points(n, 77, 4)
    scal = O.fill_fr(n, 78)
    srs = ctx.srs_register(bases)
    exp = O.g1_to_affine(O.best_multiexp(scal, bases, threads=4))
    assert np.array_equal(O.g1_to_affine(srs.msm_sharded(scal, n)), exp)
    parts = []
    for r in range(3):
        first, cnt = zk.shard_range(n, r, 3)
        parts.append(srs.msm_range(scal[first:first + cnt], first))
    assert np.array_equal(O.g1_to_affine(ctx.g1_sum(np.stack(parts))), exp)
    with pytest.raises(zk.B200zkError):
        srs.msm_range(scal[:10], n - 5)  # range runs past the bases
    srs.release()
