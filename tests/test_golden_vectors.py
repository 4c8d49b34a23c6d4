"""Committed golden vectors (tests/golden/oracle_vectors.json, made by make_oracle_vectors.py): the oracle must
reproduce them on any host (CPU test) and the CUDA path must match them bit for bit (GPU test)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

V = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_vectors.json")))


def arr(h, width):
    return np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, width).copy()


def test_oracle_reproduces_golden_vectors():
    t = V["ntt"]
    assert np.array_equal(O.best_fft(arr(t["input"], 4), arr(t["omega"], 4)[0], t["log_n"], threads=2), arr(t["output"], 4))
    d = V["domain"]
    dom = O.EvaluationDomain(d["j"], d["k"])
    coeff = dom.lagrange_to_coeff(arr(d["lagrange"], 4))
    assert np.array_equal(coeff, arr(d["coeff"], 4))
    assert np.array_equal(dom.coeff_to_extended(coeff), arr(d["extended"], 4))
    m = V["msm"]
    res = O.g1_to_affine(O.best_multiexp(arr(m["scalars"], 4), arr(m["bases"], 8), threads=3))
    assert np.array_equal(res, arr(m["result_affine"], 8)[0])
    assert O.g1_compress(res).hex() == m["result_compressed"]


@pytest.mark.gpu
def test_cuda_matches_golden_vectors(ctx, zk):
    t = V["ntt"]
    a = arr(t["input"], 4)
    ctx.best_fft(a, arr(t["omega"], 4)[0], t["log_n"])
    assert np.array_equal(a, arr(t["output"], 4))
    d = V["domain"]
    dom = zk.EvaluationDomain(ctx, d["j"], d["k"])
    col = arr(d["lagrange"], 4)
    dom.lagrange_to_coeff(col)
    assert np.array_equal(col, arr(d["coeff"], 4))
    ext = dom.coeff_to_extended(col)
    assert np.array_equal(ext, arr(d["extended"], 4))
    assert np.array_equal(dom.extended_to_coeff(ext.copy()), arr(d["back"], 4))
    m = V["msm"]
    got = ctx.best_multiexp(arr(m["scalars"], 4), arr(m["bases"], 8))
    assert np.array_equal(got[:8], arr(m["result_affine"], 8)[0])  # normalised (x, y, 1): bytes equal the affine point
