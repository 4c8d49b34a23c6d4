"""SURVEY.md §8(d) config 4 / BASELINE configs[3]: one MSM of 2^20 .. 2^26 points sharded by point range over N GPUs.

One process per GPU (torchrun); every rank holds the full SRS handle (generated on the device: [s_i]G, precomputed tables
when they fit) and calls b200zk_msm_g1_sharded with its slice of the scalars: local Pippenger over its point range, ONE
ncclAllGather of the 96-byte partial points on the context-owned communicator, local sum (csrc/comm.cu).  Timed on the device
(CUDA events around the whole call, MAX over ranks), uniform and witness-like scalars.  One JSON line per size on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/msm_sharded_scaling.py "20,22,24,26"
"""
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tools"))
zk = importlib.import_module("scroll-prover_b200")
from quick_time import rand_fr  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = zk.Context(local)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        ctx.comm_init_torch(dist)
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,22,24,26").split(",")]
    for log_n in sizes:
        n = 1 << log_n
        lo, m = zk.shard_range(n, rank, world)
        g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        ctx.g1_generator_mul_batch(rand_fr(n, 1000 * log_n), out=g)  # the same SRS on every rank
        srs = ctx.srs_register(g)
        del g
        sc = rand_fr(m, 5000 * log_n + rank)
        sel = torch.rand(m, device="cuda")
        small = torch.zeros((m, 4), dtype=torch.int64, device="cuda")
        small[:, 0] = torch.where((sel >= 0.6) & (sel < 0.9), torch.randint(0, 1 << 16, (m,), dtype=torch.int64, device="cuda"),
                                  torch.zeros(m, dtype=torch.int64, device="cuda"))
        torch.cuda.synchronize()
        wl = torch.where((sel >= 0.9).unsqueeze(1), sc, ctx.poly_scale(small, zk.fr_from_int(1 << 256))).contiguous()
        torch.cuda.synchronize()
        row = {"op": "msm_sharded", "log_n": log_n, "world": world}
        for kind, scal in (("uniform", sc), ("witness_like", wl)):
            times = []
            for it in range(4):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                srs.msm_sharded(scal, n)
                e1.record()
                torch.cuda.synchronize()
                t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if it:  # first iteration warms up
                    times.append(float(t.item()))
            st = ctx.msm_last_stats()
            best = min(times)
            row[kind] = {"ms": round(best, 3), "c": st["window_bits"], "W": st["n_windows"], "Mpoints_s": round(n / best / 1e3, 1),
                         "Gadds_s_nominal": round(n * st["n_windows"] / best / 1e6, 3)}
        if rank == 0:
            print(json.dumps(row), flush=True)
        srs.release()
        del sc, wl, small
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
