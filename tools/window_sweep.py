"""Sweep the Pippenger window width of the plain (non-precomputed) MSM path per size (development aid).

Prints one JSON line per (log_n, c) with the best-of-3 device time, so that pick_window()'s cost model in
csrc/msm.cu can be checked against measurements.  usage: window_sweep.py "14,16,18,20,22,24"
"""
import importlib
import json
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
zk = importlib.import_module("scroll-prover_b200")
sys.path.insert(0, os.path.join(_ROOT, "tools"))
from quick_time import rand_fr, timeit  # noqa: E402


def main():
    ctx = zk.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.srs_set_precompute(False)
    for log_n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "14,16,18,20,22,24").split(",")]:
        n = 1 << log_n
        g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        ctx.g1_generator_mul_batch(rand_fr(n, 100 + log_n), out=g)
        srs = ctx.srs_register(g)
        sc = rand_fr(n, 200 + log_n)
        ctx.msm_set_window(0)
        srs.msm(sc)
        auto_c = ctx.msm_last_stats()["window_bits"]
        rows = []
        for c in range(max(5, log_n - 9), min(23, log_n + 1) + 1):
            W = 254 // c + 1
            if n * W >= 4.0e9:
                continue
            ctx.msm_set_window(c)
            best, _ = timeit(lambda: srs.msm(sc), reps=3, warm=1)
            rows.append((c, best))
        ctx.msm_set_window(0)
        bc = min(rows, key=lambda r: r[1])
        print(json.dumps({"log_n": log_n, "auto_c": auto_c, "best_c": bc[0], "best_ms": round(bc[1], 4),
                          "auto_ms": round(dict(rows).get(auto_c, float("nan")), 4),
                          "sweep": {str(c): round(ms, 4) for c, ms in rows}}), flush=True)
        srs.release() if hasattr(srs, "release") else None
        del g, sc
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
