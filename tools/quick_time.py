"""Quick device timing of the hot kernels (development aid; bench.py is the judged harness)."""
import importlib
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
zk = importlib.import_module("scroll-prover_b200")

R = zk.R_MOD


def rand_fr(n, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    t = torch.randint(-(2**63), 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    t[:, 3] &= 0x0FFFFFFFFFFFFFFF  # < 2^252 < r: valid Montgomery limbs
    return t


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ctx = zk.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)  # a non-default stream so that the library launches where our events are
    ctx.set_stream(stream.cuda_stream)
    out = []
    for log_n in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "16,20,22,24").split(",")]:
        n = 1 << log_n
        a = rand_fr(n, log_n)
        w = zk.fr_from_int(pow(zk._ROOT_OF_UNITY, 1 << (28 - log_n), R))
        best, med = timeit(lambda: ctx.best_fft(a, w, log_n))
        ctx.profile_enable(True); ctx.profile_reset(); ctx.best_fft(a, w, log_n); prof = ctx.profile_read(); ctx.profile_enable(False)
        bf = n / 2 * log_n
        out.append({"op": "ntt", "log_n": log_n, "ms_best": best, "ms_med": med, "Gbutterflies_s": bf / best / 1e6,
                    "prof": {k: round(v["ms"], 4) for k, v in prof.items() if v["count"]}})
        print(json.dumps(out[-1]), flush=True)
    for log_n in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,20,22").split(",")]:
        n = 1 << log_n
        s = rand_fr(n, 100 + log_n)
        g = torch.empty((n, 8), dtype=torch.int64, device="cuda")
        t0 = time.time()
        ctx.g1_generator_mul_batch(s, out=g)
        torch.cuda.synchronize()
        gen_s = time.time() - t0
        srs = ctx.srs_register(g)
        sc = rand_fr(n, 200 + log_n)
        best, med = timeit(lambda: srs.msm(sc), reps=3, warm=1)
        st = ctx.msm_last_stats()
        ctx.profile_enable(True); ctx.profile_reset(); srs.msm(sc); prof = ctx.profile_read(); ctx.profile_enable(False)
        out.append({"op": "msm", "log_n": log_n, "ms_best": best, "ms_med": med, "c": st["window_bits"], "W": st["n_windows"],
                    "Gadds_s": n * st["n_windows"] / best / 1e6, "Mpoints_s": n / best / 1e3, "srs_gen_s": gen_s,
                    "prof": {k: round(v["ms"], 4) for k, v in prof.items() if v["count"]}})
        print(json.dumps(out[-1]), flush=True)
        # witness-like scalars (60 % zero, 30 % < 2^16 in canonical form, 10 % uniform): the realistic advice-column case
        sel = torch.rand(n, device="cuda")
        small = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
        small[:, 0] = torch.where((sel >= 0.6) & (sel < 0.9), torch.randint(0, 1 << 16, (n,), dtype=torch.int64, device="cuda"), torch.zeros(n, dtype=torch.int64, device="cuda"))
        torch.cuda.synchronize()
        wl = ctx.poly_scale(small, zk.fr_from_int(1 << 256))
        torch.cuda.synchronize()
        wl = torch.where((sel >= 0.9).unsqueeze(1), sc, wl).contiguous()
        torch.cuda.synchronize()
        bw, mw = timeit(lambda: srs.msm(wl), reps=3, warm=1)
        ctx.profile_enable(True); ctx.profile_reset(); srs.msm(wl); profw = ctx.profile_read(); ctx.profile_enable(False)
        out[-1]["witness_like_ms"] = bw
        out[-1]["witness_like_prof"] = {k: round(v["ms"], 4) for k, v in profw.items() if v["count"]}
        ctx.msm_total_adds(reset=True)
        srs.msm(wl)
        ctx.synchronize()
        adds = ctx.msm_total_adds(reset=True)  # non-zero signed digits actually accumulated
        print(json.dumps({"op": "msm_witness_like", "log_n": log_n, "ms_best": bw, "c": st["window_bits"], "W": st["n_windows"],
                          "actual_adds": adds, "Gadds_s": adds / bw / 1e6, "Mpoints_s": n / bw / 1e3,
                          "prof": out[-1]["witness_like_prof"]}), flush=True)
        srs.release()
        del g
    json.dump(out, open("gpurun_out/quick_time.json", "w"), indent=1)


if __name__ == "__main__":
    main()
