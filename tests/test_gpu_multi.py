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
    # evaluate_h sharded by ROW RANGE: every rank evaluates its slice of the extended domain with a program that reads rotated
    # columns (wrapping over the whole domain) and folds into the previous value, then one in-place all-gather over NVLink
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from quotient_programs import C_ADD, C_MUL, S_ADVICE, S_FIXED, S_INTER, S_PREV, S_Y  # noqa: E402

    log_size = 12
    size = 1 << log_size
    cols = [O.fill_fr(size, 600 + i) for i in range(3)]
    prev = O.fill_fr(size, 610)
    yv = O.fr_from_int(12345)
    prog = [(C_MUL, (S_ADVICE, 0, 1), (S_ADVICE, 1, 2), None), (C_ADD, (S_INTER, 0, 0), (S_FIXED, 0, 0), None),
            (C_MUL, (S_PREV, 0, 0), (S_Y, 0, 0), None), (C_ADD, (S_INTER, 2, 0), (S_INTER, 1, 0), None)]
    rotations = [0, 3, -5]
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    exp_vals = O.graph_evaluate(prog, e, rotations, [cols[2]], [cols[0], cols[1]], [], e, z, z, z, yv, None, prev.copy(), log_size, 4)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()
    g = ctx.graph(prog, e, rotations)
    vals = dev(prev)
    first, cnt = zk.shard_range(size, rank, world)
    torch.cuda.synchronize()
    g.evaluate(vals, log_size, 4, fixed=[dev(cols[2])], advice=[dev(cols[0]), dev(cols[1])], y=yv, rows=(first, cnt))
    ctx.synchronize()
    mine = vals.cpu().numpy().view(np.uint64)
    ok &= bool(np.array_equal(mine[first:first + cnt], exp_vals[first:first + cnt]))
    other = np.ones(size, bool)
    other[first:first + cnt] = False
    ok &= bool(np.array_equal(mine[other], prev[other]))  # rows of other ranks untouched
    ctx.allgather_rows(vals, log_size)
    ctx.synchronize()
    ok &= bool(np.array_equal(vals.cpu().numpy().view(np.uint64), exp_vals))
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


def test_graph_evaluate_rows_in_slices_equals_the_full_evaluation(ctx, zk):
    """b200zk_graph_evaluate_rows over three ragged row slices (one GPU) == b200zk_graph_evaluate == the oracle; b200zk_allgather_rows
    is a no-op without a communicator; a slice past the domain is rejected."""
    import torch

    from oracle import oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from quotient_programs import C_ADD, C_MUL, S_ADVICE, S_INTER, S_PREV, S_X

    log_size, size = 11, 1 << 11
    cols = [O.fill_fr(size, 700 + i) for i in range(2)]
    prev = O.fill_fr(size, 710)
    prog = [(C_MUL, (S_ADVICE, 0, 1), (S_ADVICE, 1, 0), None), (C_ADD, (S_INTER, 0, 0), (S_X, 0, 0), None), (C_ADD, (S_INTER, 1, 0), (S_PREV, 0, 0), None)]
    rotations = [0, -7]
    z, e = O.fr_from_int(0), np.zeros((0, 4), np.uint64)
    w = O.fr_from_int(pow(zk._ROOT_OF_UNITY, 1 << (28 - log_size), zk.R_MOD))
    exp = O.graph_evaluate(prog, e, rotations, [], cols, [], e, z, z, z, z, w, prev.copy(), log_size, 2)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()
    g = ctx.graph(prog, e, rotations)
    dcols = [dev(c) for c in cols]
    vals = dev(prev)
    torch.cuda.synchronize()
    for first, cnt in ((0, 700), (700, 1), (701, size - 701)):
        g.evaluate(vals, log_size, 2, advice=dcols, extended_omega=w, rows=(first, cnt))
    ctx.allgather_rows(vals, log_size)
    ctx.synchronize()
    assert np.array_equal(vals.cpu().numpy().view(np.uint64), exp)
    with pytest.raises(zk.B200zkError):
        g.evaluate(vals, log_size, 2, advice=dcols, extended_omega=w, rows=(size - 3, 4))
