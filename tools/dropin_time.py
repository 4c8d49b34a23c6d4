"""What the CHEAP integration buys (each halo2_proofs function body replaced by one FFI call on host slices,
no session): per-call wall time including the host<->device copies of every vector.

  best_fft        host slice in, host slice out (in place)            b200zk_ntt_fr on a host pointer
  best_multiexp   host scalars + host bases, bases uploaded per call   b200zk_msm_g1_bases
  commit          host scalars, bases resident (ParamsKZG handle)      b200zk_msm_g1
for pageable and pinned host memory, beside the device-resident kernel time of the same operation.
usage: dropin_time.py "20,24,26" "20,24"      -> JSON lines (profiles/dropin_r02.jsonl)"""
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tools"))
zk = importlib.import_module("scroll-prover_b200")
from quick_time import rand_fr  # noqa: E402

R = zk.R_MOD


def best(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


def main():
    ctx = zk.Context(0)
    ntt_sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "20,24").split(",")]
    msm_sizes = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "20,24").split(",")]
    for log_n in ntt_sizes:
        n = 1 << log_n
        w = zk.fr_from_int(pow(zk._ROOT_OF_UNITY, 1 << (28 - log_n), R))
        dev = rand_fr(n, log_n)
        pageable = dev.cpu().numpy().view(np.uint64)
        pinned = dev.cpu().pin_memory()
        ctx.best_fft(dev, w, log_n)  # warm: twiddle table
        row = {"op": "best_fft", "log_n": log_n, "bytes_each_way": 32 * n,
               "device_resident_ms": round(best(lambda: ctx.best_fft(dev, w, log_n)), 3),
               "host_pinned_in_out_ms": round(best(lambda: ctx.best_fft(pinned, w, log_n)), 3),
               "host_pageable_in_out_ms": round(best(lambda: ctx.best_fft(pageable, w, log_n)), 3)}
        print(json.dumps(row), flush=True)
        del dev, pageable, pinned
    for log_n in msm_sizes:
        n = 1 << log_n
        g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        ctx.g1_generator_mul_batch(rand_fr(n, 100 + log_n), out=g)
        sc = rand_fr(n, 200 + log_n)
        g_host, sc_host = g.cpu().numpy().view(np.uint64), sc.cpu().numpy().view(np.uint64)
        g_pin, sc_pin = g.cpu().pin_memory(), sc.cpu().pin_memory()
        srs = ctx.srs_register(g)
        srs.msm(sc)
        row = {"op": "best_multiexp", "log_n": log_n, "scalar_bytes": 32 * n, "base_bytes": 64 * n,
               "device_resident_ms": round(best(lambda: srs.msm(sc)), 3),
               "commit_host_pinned_scalars_ms": round(best(lambda: srs.msm(sc_pin)), 3),
               "commit_host_pageable_scalars_ms": round(best(lambda: srs.msm(sc_host)), 3),
               "best_multiexp_host_pinned_bases_and_scalars_ms": round(best(lambda: ctx.best_multiexp(sc_pin, g_pin)), 3),
               "best_multiexp_host_pageable_bases_and_scalars_ms": round(best(lambda: ctx.best_multiexp(sc_host, g_host)), 3)}
        print(json.dumps(row), flush=True)
        srs.release()
        del g, sc, g_host, sc_host, g_pin, sc_pin
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
