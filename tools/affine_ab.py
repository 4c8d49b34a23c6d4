"""A/B of the experimental batched-affine bucket accumulation (B200ZK_MSM_AFFINE=1) against the default XYZZ path.

First measurement of the next round (DESIGN.md "Measured leads"): for every size, uniform and witness-like scalars, the
two paths must return the same point (checked first), then both are timed with their per-phase breakdown.
usage: affine_ab.py "16,20,22,24"
"""
import importlib
import json
import os
import sys

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tools"))
zk = importlib.import_module("scroll-prover_b200")
from quick_time import rand_fr, timeit  # noqa: E402


def witness_like(ctx, sc, n):
    sel = torch.rand(n, device="cuda")
    small = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    small[:, 0] = torch.where((sel >= 0.6) & (sel < 0.9), torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device="cuda"),
                              torch.zeros(n, dtype=torch.int64, device="cuda"))
    torch.cuda.synchronize()
    wl = ctx.poly_scale(small, zk.fr_from_int(1 << 256))  # to Montgomery form
    torch.cuda.synchronize()
    return torch.where((sel >= 0.9).unsqueeze(1), sc, wl).contiguous()


def make_ctx(affine: bool):
    os.environ["B200ZK_MSM_AFFINE"] = "1" if affine else "0"  # read once per context
    ctx = zk.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    return ctx


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,20,22,24").split(",")]
    for log_n in sizes:
        n = 1 << log_n
        row = {"log_n": log_n}
        results = {}
        for name, affine in (("xyzz", False), ("affine", True)):
            ctx = make_ctx(affine)
            g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
            ctx.g1_generator_mul_batch(rand_fr(n, 100 + log_n), out=g)
            srs = ctx.srs_register(g)
            sc = rand_fr(n, 200 + log_n)
            torch.manual_seed(log_n)
            wl = witness_like(ctx, sc, n)
            for kind, scal in (("uniform", sc), ("witness", wl)):
                res = srs.msm(scal)
                results[(name, kind)] = np.array(res)
                best, _ = timeit(lambda: srs.msm(scal), reps=3, warm=1)
                ctx.profile_enable(True); ctx.profile_reset(); srs.msm(scal); prof = ctx.profile_read(); ctx.profile_enable(False)
                row[f"{name}_{kind}_ms"] = round(best, 3)
                row[f"{name}_{kind}_prof"] = {k: round(v["ms"], 3) for k, v in prof.items() if v["count"]}
            st = ctx.msm_last_stats()
            row["c"], row["W"] = st["window_bits"], st["n_windows"]
            srs.release()
            ctx.close()
            del g, sc, wl
            torch.cuda.empty_cache()
        for kind in ("uniform", "witness"):
            row[f"{kind}_equal"] = bool(np.array_equal(results[("xyzz", kind)], results[("affine", kind)]))
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
