"""Device timing of the quotient-construction kernels (development aid; results are copied into profiles/).

usage: quotient_time.py [k]     (default 22: columns of 2^k rows, extended domain 2^(k+2))
"""
import importlib
import json
import os
import sys

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _ROOT)
sys.path.insert(0, os.path.join(_ROOT, "tools"))
sys.path.insert(0, os.path.join(_ROOT, "tests"))
zk = importlib.import_module("scroll-prover_b200")
from quick_time import rand_fr, timeit  # noqa: E402
from quotient_programs import (C_ADD, C_HORNER, C_MUL, C_SUB, DELTA, S_ADVICE, S_CONST, S_FIXED, S_INTER, S_PREV, S_Y, omega_of)  # noqa: E402


def gate_program(n_gates: int, n_advice: int):
    """n_gates custom gates of the halo2-lib flavour q * (a + b*c - d) over rotations 0..3, folded with y (Horner)."""
    calcs = []
    for g in range(n_gates):
        col = g % n_advice
        base = len(calcs)
        calcs.append((C_MUL, (S_ADVICE, col, 1), (S_ADVICE, col, 2), None))
        calcs.append((C_ADD, (S_INTER, base, 0), (S_ADVICE, col, 0), None))
        calcs.append((C_SUB, (S_INTER, base + 1, 0), (S_ADVICE, col, 3), None))
        calcs.append((C_MUL, (S_INTER, base + 2, 0), (S_FIXED, g % 2, 0), None))
    parts = [(S_INTER, 4 * g + 3, 0) for g in range(n_gates)]
    calcs.append((C_HORNER, (S_PREV, 0, 0), (S_Y, 0, 0), parts))
    return calcs


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 22
    ek = k + 2
    ctx = zk.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    fr = zk.fr_from_int
    n, size = 1 << k, 1 << ek
    # ---- GraphEvaluator over the extended domain
    n_advice, n_gates = 4, 16
    adv = [rand_fr(size, 10 + i) for i in range(n_advice)]
    fix = [rand_fr(size, 20 + i) for i in range(2)]
    vals = rand_fr(size, 30)
    calcs = gate_program(n_gates, n_advice)
    g = ctx.graph(calcs, [fr(1)], [0, 1, 2, 3])
    info = g.info()
    y = fr(0x1234567)
    best, med = timeit(lambda: g.evaluate(vals, ek, 4, fixed=fix, advice=adv, y=y), reps=5, warm=2)
    muls = 3 * n_gates + n_gates  # 2 MUL + ... per gate: mul, mul, + one MAD per gate in the Horner
    col_reads = 5 * n_gates + 1
    print(json.dumps({"op": "graph_evaluate", "log_size": ek, "n_calcs": len(calcs), **info, "ms_best": best, "ms_med": med,
                      "Grows_s": size / best / 1e6, "Gmul_s": size * muls / best / 1e6,
                      "column_read_GBs": size * 32 * col_reads / best / 1e6,
                      "unique_GBs": size * 32 * (n_advice + 2 + 2) / best / 1e6}), flush=True)
    del adv, fix, vals
    torch.cuda.empty_cache()
    # ---- scans
    a = rand_fr(n, 40)
    for op, name in ((0, "prefix_product"), (1, "prefix_sum")):
        out = torch.empty_like(a)
        best, med = timeit(lambda: ctx.prefix_scan(op, a, fr(3), out=out), reps=5, warm=2)
        print(json.dumps({"op": name, "log_n": k, "ms_best": best, "ms_med": med, "GBs": n * 64 * 1.5 / best / 1e6}), flush=True)
    # ---- permutation product, 3 columns (cs_degree - 2 = 3 for the degree-5 systems of the layer configs)
    v = [rand_fr(n, 50 + i) for i in range(3)]
    s = [rand_fr(n, 60 + i) for i in range(3)]
    z = torch.empty_like(a)
    w = fr(omega_of(k))
    best, med = timeit(lambda: ctx.permutation_product(v, s, fr(5), fr(7), fr(1), fr(DELTA), w, k, fr(1), z), reps=5, warm=2)
    print(json.dumps({"op": "permutation_product", "k": k, "n_cols": 3, "ms_best": best, "ms_med": med,
                      "Melem_s": n / best / 1e3}), flush=True)
    # ---- log-derivative running sum, 2 inputs
    f = [rand_fr(n, 70 + i) for i in range(2)]
    t, m = rand_fr(n, 80), rand_fr(n, 81)
    best, med = timeit(lambda: ctx.logup_running_sum(f, t, m, fr(5), k, fr(0), z), reps=5, warm=2)
    print(json.dumps({"op": "logup_running_sum", "k": k, "n_inputs": 2, "ms_best": best, "ms_med": med,
                      "Melem_s": n / best / 1e3}), flush=True)
    # ---- batch inversion alone, for reference
    b = rand_fr(n, 90)
    best, med = timeit(lambda: ctx.batch_invert(b), reps=5, warm=2)
    print(json.dumps({"op": "batch_invert", "log_n": k, "ms_best": best, "ms_med": med, "Melem_s": n / best / 1e3}), flush=True)


if __name__ == "__main__":
    main()
