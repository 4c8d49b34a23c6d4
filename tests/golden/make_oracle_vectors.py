"""Generates tests/golden/oracle_vectors.json: small seeded input/output vectors of the hot path computed by the
CPU oracle (and cross-checked against the independent big-int model when written).  Both the oracle (CPU test)
and the CUDA path (GPU test) are compared against these committed bytes.
    python tests/golden/make_oracle_vectors.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from oracle import pyref as P  # noqa: E402


def hx(a):
    return np.ascontiguousarray(a, np.uint64).tobytes().hex()


def main():
    out = {}
    log_n = 6
    a = O.fill_fr(1 << log_n, 0xA11CE)
    w = O.fr_from_int(P.omega_for(log_n))
    f = O.best_fft(a, w, log_n, threads=1)
    assert O.frs_to_ints(f) == P.dft(O.frs_to_ints(a), P.omega_for(log_n))
    out["ntt"] = {"log_n": log_n, "omega": hx(w), "input": hx(a), "output": hx(f)}
    k = 4
    dom = O.EvaluationDomain(5, k)
    col = O.fill_fr(1 << k, 0xB0B, witness_like=True)
    coeff = dom.lagrange_to_coeff(col)
    ext = dom.coeff_to_extended(coeff)
    assert O.frs_to_ints(ext) == P.coeff_to_extended(O.frs_to_ints(coeff), k, k + 2)
    out["domain"] = {"j": 5, "k": k, "lagrange": hx(col), "coeff": hx(coeff), "extended": hx(ext),
                     "back": hx(dom.extended_to_coeff(ext))}
    n = 33
    bases = O.fill_points(n, 0xCAFE, 4)
    sc = O.fill_fr(n, 0xD00D)
    sc[0] = 0
    sc[1] = O.fr_from_int(O.R_MOD - 1)
    res = O.g1_to_affine(O.best_multiexp(sc, bases, threads=1))
    py = P.msm(O.frs_to_ints(sc), [(O.fq_to_int(b[:4]), O.fq_to_int(b[4:])) for b in bases])
    assert (O.fq_to_int(res[:4]), O.fq_to_int(res[4:])) == py
    out["msm"] = {"n": n, "scalars": hx(sc), "bases": hx(bases), "result_affine": hx(res),
                  "result_compressed": O.g1_compress(res).hex()}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.json"), "w"), indent=1)
    print("ok")


if __name__ == "__main__":
    main()
