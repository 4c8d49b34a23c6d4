"""Sweep the accumulate chunk length (B200ZK_ACC_L: sorted entries per thread of msm_accumulate) per MSM size.

Development aid for the small-MSM latency floor documented in DESIGN.md ("Measured leads"): with the default 256 entries
per thread an MSM of N*W < ~20 M entries does not fill the machine and sits on the serial chain of 256 mixed additions.
Prints one JSON line per (log_n, path) with the time at every chunk length.  usage: acc_l_sweep.py "12,14,16,18,20,22"
"""
import importlib
import json
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tools"))
zk = importlib.import_module("scroll-prover_b200")
from quick_time import rand_fr, timeit  # noqa: E402


def main():
    sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "12,14,16,18,20,22").split(",")]
    for pre in (0, 1):
        for log_n in sizes:
            n = 1 << log_n
            row = {"precompute": pre, "log_n": log_n, "ms": {}}
            for L in (16, 32, 64, 128, 256):
                os.environ["B200ZK_ACC_L"] = str(L)  # read once per context
                ctx = zk.Context(0)
                stream = torch.cuda.Stream()
                torch.cuda.set_stream(stream)
                ctx.set_stream(stream.cuda_stream)
                ctx.srs_set_precompute(bool(pre))
                g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
                ctx.g1_generator_mul_batch(rand_fr(n, 100 + log_n), out=g)
                srs = ctx.srs_register(g)
                sc = rand_fr(n, 200 + log_n)
                best, _ = timeit(lambda: srs.msm(sc), reps=3, warm=1)
                st = ctx.msm_last_stats()
                row["c"], row["W"] = st["window_bits"], st["n_windows"]
                row["ms"][str(L)] = round(best, 4)
                srs.release()
                ctx.close()
                del g, sc
                torch.cuda.empty_cache()
            row["best_L"] = min(row["ms"], key=lambda k: row["ms"][k])
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
