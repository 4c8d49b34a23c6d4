"""Per-phase device time of small MSMs on both paths (plain bases / precomputed tables): where the fixed cost goes."""
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

ctx = zk.Context(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
for pre in (0, 1):
    ctx.srs_set_precompute(bool(pre))
    for log_n in (10, 14, 16, 18):
        n = 1 << log_n
        g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        ctx.g1_generator_mul_batch(rand_fr(n, 100 + log_n), out=g)
        srs = ctx.srs_register(g)
        sc = rand_fr(n, 200 + log_n)
        best, med = timeit(lambda: srs.msm(sc), reps=3, warm=1)
        st = ctx.msm_last_stats()
        ctx.profile_enable(True); ctx.profile_reset(); srs.msm(sc); prof = ctx.profile_read(); ctx.profile_enable(False)
        print(json.dumps({"precompute": pre, "log_n": log_n, "c": st["window_bits"], "W": st["n_windows"], "ms_best": round(best, 4),
                          "prof": {k: round(v["ms"], 4) for k, v in prof.items() if v["count"]}}), flush=True)
        srs.release()
