"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the point-range MSM sharding + all-gather +
sum path of scroll-prover_b200/multi.py (with the oracle standing in for the per-rank device MSM) and the
job fan-out used by bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import importlib

    import torch.distributed as dist

    from oracle import oracle as O

    multi = importlib.import_module("scroll-prover_b200.multi")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bases = O.fill_points(n, 31, 2)
    scal = O.fill_fr(n, 32, witness_like=True)
    lo, hi = multi.shard_range(n, rank, world)

    def local_msm(s):
        return O.best_multiexp(s, bases[lo:hi], threads=1)

    def g1_sum(pts):
        acc = pts[0]
        for p in pts[1:]:
            acc = O.g1_add(acc, p)
        return acc

    total = multi.msm_sharded(local_msm, g1_sum, scal[lo:hi], dist)
    exp = O.best_multiexp(scal, bases, threads=1)
    q.put((rank, bool(np.array_equal(O.g1_to_affine(total), O.g1_to_affine(exp)))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [257])
def test_point_range_sharded_msm_world2(n):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_partitions():
    sys.path.insert(0, ROOT)
    import importlib

    multi = importlib.import_module("scroll-prover_b200.multi")
    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 4, 8):
            rs = [multi.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1


def test_job_fanout_is_balanced_and_complete():
    """bench.py --gpus N: every layer of the chunk step is fanned out on its own (the layers run in sequence)"""
    sys.path.insert(0, ROOT)
    import bench

    for layer in bench.CHUNK_LAYERS:
        jobs = bench.make_jobs(layer)
        n_msm, n_ntt = layer.wit + layer.uni + layer.coeff, layer.wit + layer.uni
        assert sum(1 for j in jobs if j[0] in ("lmsm", "msm")) == n_msm and sum(1 for j in jobs if j[0] == "ntt") == n_ntt
        biggest = max(j[2] for j in jobs)
        for world in (1, 2, 4, 8):
            parts = bench.assign_jobs(jobs, world)
            assert sum(len(p) for p in parts) == len(jobs)
            loads = [sum(j[2] for j in p) for p in parts]
            assert max(loads) <= sum(loads) / world + biggest  # within one job of perfect balance


def test_shard_ranges_partition_and_lpt_balances():
    """host logic properties: shard_range tiles [0, n) without gaps for every world size; assign_jobs (LPT) places every job
    once and its makespan is within 4/3 of the trivial lower bound"""
    import importlib
    import random

    multi = importlib.import_module("scroll-prover_b200.multi")
    rng = random.Random(5)
    for n in (0, 1, 7, 8, 1000, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            ranges = [multi.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1
    for trial in range(50):
        world = rng.choice([1, 2, 4, 8])
        costs = [rng.choice([11, 18, 44, 16]) * rng.uniform(0.9, 1.1) for _ in range(rng.randrange(1, 80))]
        plan = multi.assign_jobs(costs, world)
        assert sorted(i for r in plan for i in r) == list(range(len(costs)))
        loads = [sum(costs[i] for i in r) for r in plan]
        lower = max(max(costs), sum(costs) / world)
        assert max(loads) <= 4 / 3 * lower + 1e-9


def test_abi_shard_range_tiles_the_points_and_matches_the_host_helper():
    """b200zk_shard_range (no device needed) is the partition b200zk_msm_g1_sharded uses: contiguous, covering,
    sizes differing by at most one, identical to multi.shard_range."""
    import importlib

    zk = importlib.import_module("scroll-prover_b200")
    multi = importlib.import_module("scroll-prover_b200.multi")
    for n in (0, 1, 5, 8, 1000, (1 << 17) + 3, 1 << 26):
        for world in (1, 2, 3, 4, 8):
            nxt = 0
            sizes = []
            for r in range(world):
                first, cnt = zk.shard_range(n, r, world)
                assert first == nxt and (first, first + cnt) == multi.shard_range(n, r, world)
                nxt += cnt
                sizes.append(cnt)
            assert nxt == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(zk.B200zkError):
        zk.shard_range(10, 3, 3)
