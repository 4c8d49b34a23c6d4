"""EXACT parity at BASELINE sizes (slow, -m gpu): the CUDA path against the live oracle (all host threads) AND against the
committed SHA-256 digests of the oracle's outputs (tests/golden/big_digests.json, written by tests/golden/make_big_digests.py):
best_multiexp at 2^22 / 2^24 for uniform and witness-like scalars, lagrange_to_coeff at 2^24, coeff_to_extended at 2^26
(every element, not sampled positions), g_to_lagrange at 2^14 / 2^16.  Reference functions: halo2_proofs/src/arithmetic.rs,
src/poly/domain.rs, src/poly/kzg/commitment.rs @ e5ddf67 (oracle/halo2_arith.c, halo2_domain.c, halo2_params.c)."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_big_digests as G  # noqa: E402  (the input recipes live beside the digests)

pytestmark = pytest.mark.gpu
DIGESTS = json.load(open(os.path.join(ROOT, "tests", "golden", "big_digests.json")))
THREADS = os.cpu_count() or 8


@pytest.mark.parametrize("log_n", [22, 24])
@pytest.mark.parametrize("witness_like", [False, True])
def test_best_multiexp_exact_at_baseline_sizes(ctx, log_n, witness_like):
    bases, scal, exp = G.msm_case(log_n, witness_like, THREADS)
    srs = ctx.srs_register(bases)  # precomputed tables (one bucket set), the production path
    got = O.g1_to_affine(srs.msm(scal))
    srs.release()
    assert np.array_equal(got, exp)
    assert G.sha(got) == DIGESTS[f"msm_{log_n}_{'witness' if witness_like else 'uniform'}"]
    # and the plain per-window path on the same inputs
    ctx.srs_set_precompute(False)
    try:
        srs = ctx.srs_register(bases)
    finally:
        ctx.srs_set_precompute(True)
    assert np.array_equal(O.g1_to_affine(srs.msm(scal)), exp)
    srs.release()


def test_transforms_exact_at_degree_24(ctx, zk):
    """every element of lagrange_to_coeff(2^24) and coeff_to_extended(2^26) -- the degree-24 layer's transforms"""
    k = 24
    a = G.ntt_inputs(k)
    dom, dom_o = zk.EvaluationDomain(ctx, 5, k), O.EvaluationDomain(5, k)
    exp_c = dom_o.lagrange_to_coeff(a, THREADS)
    got_c = a.copy()
    dom.lagrange_to_coeff(got_c)
    assert np.array_equal(got_c, exp_c)
    assert G.sha(got_c) == DIGESTS["lagrange_to_coeff_24"]
    got_e = np.asarray(dom.coeff_to_extended(got_c))
    assert G.sha(got_e) == DIGESTS["coeff_to_extended_26"]
    exp_e = dom_o.coeff_to_extended(exp_c, THREADS)
    assert np.array_equal(got_e, exp_e)
    # and back: extended_to_coeff of the exact coset evaluations returns the coefficients (zero above degree 2^24)
    back = np.asarray(dom.extended_to_coeff(got_e.copy()))
    assert np.array_equal(back[: 1 << k], exp_c) and not back[1 << k:].any()


@pytest.mark.parametrize("k", [14, 16])
def test_g_to_lagrange_exact(ctx, k):
    """Params::downsize's G1 inverse FFT at sizes where the oracle's G1 FFT is still affordable"""
    g = O.fill_points_chain(1 << k, 9300 + k, G.GEN_THREADS)
    got = ctx.g_to_lagrange(g, k)
    assert G.sha(got) == DIGESTS[f"g_to_lagrange_{k}"]
    if k <= 14:
        assert np.array_equal(got, O.g_to_lagrange(g, k, THREADS))
