"""Times Params::downsize's device part (b200zk_g_to_lagrange = inverse G1 FFT + 1/n) for the given degrees.
usage: g1fft_time.py "16,18,20,21"   ->  one JSON line per k (profiles/g1fft_r02.jsonl)"""
import importlib
import json
import os
import sys
import time

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tools"))
zk = importlib.import_module("scroll-prover_b200")
from quick_time import rand_fr  # noqa: E402


def main():
    ctx = zk.Context(0)
    for k in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,18,20").split(",")]:
        n = 1 << k
        g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        ctx.g1_generator_mul_batch(rand_fr(n, 7000 + k), out=g)
        out = torch.empty_like(g)
        torch.cuda.synchronize()
        ts = []
        for _ in range(2):
            t0 = time.perf_counter()
            ctx.g_to_lagrange(g, k, out=out)
            ctx.synchronize()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        bf = n // 2 * k
        print(json.dumps({"op": "g_to_lagrange", "k": k, "s_best": round(min(ts), 4), "G1_butterflies": bf,
                          "M_butterflies_per_s": round(bf / min(ts) / 1e6, 2)}), flush=True)
        del g, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
