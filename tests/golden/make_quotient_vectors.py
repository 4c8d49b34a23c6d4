"""Generates tests/golden/quotient_vectors.json: small seeded input/output vectors of the quotient-construction steps
(GraphEvaluator program, permutation z product, log-derivative running sum, the evaluate_h permutation section) computed by
the CPU oracle and cross-checked against the big-integer model when written.  tests/test_quotient_golden.py holds the oracle to
these committed bytes (so the restatement cannot drift silently between rounds).
    python tests/golden/make_quotient_vectors.py
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from quotient_programs import (DELTA, R_MOD, model_evaluate, model_logup, model_permutation_product, omega_of, random_program)  # noqa: E402
from h_terms_programs import permutation_terms_program  # noqa: E402


def hx(a):
    return np.ascontiguousarray(a, np.uint64).tobytes().hex()


def main():
    out = {}
    rng = random.Random(0x90_1D)
    # ---- a 40-calculation program over 2 fixed / 3 advice / 1 instance columns, 16 rows, rot_scale 2
    log_size, rot_scale = 4, 2
    size = 1 << log_size
    calcs, constants, rotations = random_program(0x5EED, 40, 2, 3, 1, 2, 5)
    cols = lambda c: [[rng.randrange(R_MOD) for _ in range(size)] for _ in range(c)]
    fx, ad, ins = cols(2), cols(3), cols(1)
    ch = [rng.randrange(R_MOD) for _ in range(2)]
    beta, gamma, theta, y = (rng.randrange(R_MOD) for _ in range(4))
    prev = [rng.randrange(R_MOD) for _ in range(size)]
    w = omega_of(log_size)
    F = O.frs_from_ints
    got = O.graph_evaluate(calcs, F(constants), rotations, [F(c) for c in fx], [F(c) for c in ad], [F(c) for c in ins], F(ch),
                           O.fr_from_int(beta), O.fr_from_int(gamma), O.fr_from_int(theta), O.fr_from_int(y), O.fr_from_int(w), F(prev),
                           log_size, rot_scale)
    assert O.frs_to_ints(got) == model_evaluate(calcs, constants, rotations, fx, ad, ins, ch, beta, gamma, theta, y, w, prev, log_size, rot_scale)
    out["graph"] = {"calcs": [[op, list(a), list(b) if b else None, [list(p) for p in ps] if ps else None] for op, a, b, ps in calcs],
                    "constants": hx(F(constants)), "rotations": rotations, "fixed": [hx(F(c)) for c in fx], "advice": [hx(F(c)) for c in ad],
                    "instance": [hx(F(c)) for c in ins], "challenges": hx(F(ch)), "beta": hx(O.fr_from_int(beta)),
                    "gamma": hx(O.fr_from_int(gamma)), "theta": hx(O.fr_from_int(theta)), "y": hx(O.fr_from_int(y)),
                    "extended_omega": hx(O.fr_from_int(w)), "previous": hx(F(prev)), "log_size": log_size, "rot_scale": rot_scale,
                    "output": hx(got)}
    # ---- permutation product, 3 columns, k = 4
    k = 4
    n = 1 << k
    v, s = [[rng.randrange(R_MOD) for _ in range(n)] for _ in range(3)], [[rng.randrange(R_MOD) for _ in range(n)] for _ in range(3)]
    z0, dws = rng.randrange(R_MOD), pow(DELTA, 3, R_MOD)
    z = O.permutation_product([F(c) for c in v], [F(c) for c in s], O.fr_from_int(beta), O.fr_from_int(gamma), O.fr_from_int(dws),
                              O.fr_from_int(DELTA), O.fr_from_int(omega_of(k)), k, O.fr_from_int(z0))
    assert O.frs_to_ints(z) == model_permutation_product(v, s, beta, gamma, dws, DELTA, omega_of(k), k, z0)
    out["permutation_product"] = {"k": k, "values": [hx(F(c)) for c in v], "sigma": [hx(F(c)) for c in s], "beta": hx(O.fr_from_int(beta)),
                                  "gamma": hx(O.fr_from_int(gamma)), "delta_omega_start": hx(O.fr_from_int(dws)),
                                  "delta": hx(O.fr_from_int(DELTA)), "omega": hx(O.fr_from_int(omega_of(k))), "z_init": hx(O.fr_from_int(z0)),
                                  "z": hx(z)}
    # ---- log-derivative running sum, 2 inputs
    f = [[rng.randrange(R_MOD) for _ in range(n)] for _ in range(2)]
    t = [rng.randrange(R_MOD) for _ in range(n)]
    m = [rng.randrange(0, 4) for _ in range(n)]
    p0 = rng.randrange(R_MOD)
    phi = O.logup_running_sum([F(c) for c in f], F(t), F(m), O.fr_from_int(beta), k, O.fr_from_int(p0))
    assert O.frs_to_ints(phi) == model_logup(f, t, m, beta, k, p0)
    out["logup"] = {"k": k, "inputs": [hx(F(c)) for c in f], "table": hx(F(t)), "m": hx(F(m)), "beta": hx(O.fr_from_int(beta)),
                    "phi_init": hx(O.fr_from_int(p0)), "phi": hx(phi)}
    # ---- evaluate_h permutation section: 2 sets, 4 columns, extended domain 2^5 (k = 3, extension 4x), last rotation -3
    k, ext = 3, 2
    ek = k + ext
    esize = 1 << ek
    col = lambda: [rng.randrange(R_MOD) for _ in range(esize)]
    zc, vc, sc = [col() for _ in range(2)], [col() for _ in range(4)], [col() for _ in range(4)]
    l0, ll, la, pv = col(), col(), col(), col()
    we = omega_of(ek)
    want = O.permutation_h_terms([F(c) for c in zc], 2, [F(c) for c in vc], [F(c) for c in sc], F(l0), F(ll), F(la), O.fr_from_int(beta),
                                 O.fr_from_int(gamma), O.fr_from_int(y), O.fr_from_int(DELTA), O.fr_from_int(we), -3, F(pv), ek, 1 << ext)
    pc, pconst, prot = permutation_terms_program(2, 2, 4, -3)
    via_program = O.graph_evaluate(pc, F(pconst), prot, [F(c) for c in sc] + [F(l0), F(ll), F(la)], [F(c) for c in zc] + [F(c) for c in vc], [],
                                   np.zeros((0, 4), np.uint64), O.fr_from_int(beta), O.fr_from_int(gamma), O.fr_from_int(0), O.fr_from_int(y),
                                   O.fr_from_int(we), F(pv), ek, 1 << ext)
    assert np.array_equal(want, via_program)
    out["permutation_h_terms"] = {"k": k, "extended_k": ek, "chunk_len": 2, "last_rotation": -3, "z": [hx(F(c)) for c in zc],
                                  "values": [hx(F(c)) for c in vc], "sigma": [hx(F(c)) for c in sc], "l0": hx(F(l0)), "l_last": hx(F(ll)),
                                  "l_active_row": hx(F(la)), "previous": hx(F(pv)), "beta": hx(O.fr_from_int(beta)),
                                  "gamma": hx(O.fr_from_int(gamma)), "y": hx(O.fr_from_int(y)), "delta": hx(O.fr_from_int(DELTA)),
                                  "extended_omega": hx(O.fr_from_int(we)), "output": hx(want)}
    json.dump(out, open(os.path.join(HERE, "quotient_vectors.json"), "w"), indent=1)
    print("ok")


if __name__ == "__main__":
    main()
