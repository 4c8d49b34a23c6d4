"""C oracle vs the independent pure-Python big-int model (oracle/pyref.py) and algebraic identities.

The reference ships no direct MSM/NTT input->output vectors, so the restated
algorithms are pinned by (a) an independent implementation and (b) uniqueness identities.
"""
import random

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P

rng = random.Random(0xB200)


def rand_fr():
    return rng.randrange(P.R_MOD)


def aff_to_py(a):
    a = np.asarray(a)
    if not a.any():
        return None
    return (O.fq_to_int(a[:4]), O.fq_to_int(a[4:]))


def py_to_aff(p):
    if p is None:
        return np.zeros(8, np.uint64)
    return np.concatenate([O.fq_from_int(p[0]), O.fq_from_int(p[1])])


def jac_to_py(j):
    return aff_to_py(O.g1_to_affine(j))


EDGE = [0, 1, 2, P.R_MOD - 1, P.R_MOD - 2, (1 << 256) % P.R_MOD, (1 << 128), (1 << 253)]


def test_field_ops_vs_bigint():
    vals = EDGE + [rand_fr() for _ in range(40)]
    for a in vals:
        for b in vals[:12]:
            A, B = O.fr_from_int(a), O.fr_from_int(b)
            assert O.fr_to_int(O.fr_mul(A, B)) == a * b % P.R_MOD
            assert O.fr_to_int(O.fr_add(A, B)) == (a + b) % P.R_MOD
            assert O.fr_to_int(O.fr_sub(A, B)) == (a - b) % P.R_MOD
        assert int.from_bytes(O.fr_to_repr(O.fr_from_int(a)), "little") == a
        if a:
            assert O.fr_to_int(O.fr_inv(O.fr_from_int(a))) == pow(a, -1, P.R_MOD)
    for a in [0, 1, P.Q_MOD - 1] + [rng.randrange(P.Q_MOD) for _ in range(20)]:
        for b in [0, 1, P.Q_MOD - 1] + [rng.randrange(P.Q_MOD) for _ in range(5)]:
            assert O.fq_to_int(O.fq_mul(O.fq_from_int(a), O.fq_from_int(b))) == a * b % P.Q_MOD
            assert O.fq_to_int(O.fq_sub(O.fq_from_int(a), O.fq_from_int(b))) == (a - b) % P.Q_MOD


def test_batch_invert_skips_zeros():
    vals = [rand_fr() for _ in range(33)]
    vals[0] = 0
    vals[17] = 0
    out = O.frs_to_ints(O.fr_batch_invert(O.frs_from_ints(vals)))
    assert out == [pow(v, -1, P.R_MOD) if v else 0 for v in vals]


def test_curve_ops_vs_affine_model():
    G = P.G1_GEN
    pts = [None, G, P.g1_mul(G, 2), P.g1_mul(G, 3), P.g1_neg(G)] + [P.g1_mul(G, rand_fr()) for _ in range(6)]
    for a in pts:
        ja = O.g1_from_affine(py_to_aff(a))
        assert jac_to_py(O.g1_double(ja)) == P.g1_add(a, a)
        for b in pts:
            jb = O.g1_from_affine(py_to_aff(b))
            exp = P.g1_add(a, b)
            assert jac_to_py(O.g1_add(ja, jb)) == exp
            assert jac_to_py(O.g1_add_mixed(ja, py_to_aff(b))) == exp
            # non-trivial Z on the left operand
            ja2 = O.g1_add(O.g1_double(ja), O.g1_from_affine(py_to_aff(P.g1_neg(a))))
            assert jac_to_py(O.g1_add_mixed(ja2, py_to_aff(b))) == exp
            assert jac_to_py(O.g1_add(ja2, jb)) == exp
    for s in [0, 1, 2, P.R_MOD - 1, rand_fr()]:
        assert jac_to_py(O.g1_mul(O.g1_from_affine(py_to_aff(G)), O.fr_from_int(s))) == P.g1_mul(G, s)
    gen = O.g1_generator()
    assert aff_to_py(gen) == G and O.g1_affine_is_on_curve(gen)
    assert O.g1_compress(np.zeros(8, np.uint64)) == P.compress(None)


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 5, 7])
@pytest.mark.parametrize("threads", [1, 4])
def test_best_fft_is_the_dft(log_n, threads):
    n = 1 << log_n
    a = [rand_fr() for _ in range(n)]
    a[0] = 0
    w = P.omega_for(log_n)
    got = O.frs_to_ints(O.best_fft(O.frs_from_ints(a), O.fr_from_int(w), log_n, threads))
    assert got == P.dft(a, w)


def test_best_fft_iterative_and_recursive_agree_large():
    log_n = 12
    a = O.fill_fr(1 << log_n, 7)
    w = O.fr_from_int(P.omega_for(log_n))
    x1 = O.best_fft(a, w, log_n, threads=1)  # log_n > log_threads: recursive
    x2 = O.best_fft(a, w, log_n, threads=8)
    x3 = O.best_fft(a, w, log_n, threads=1 << 13)  # forces the iterative branch (log_n <= log_threads)
    assert np.array_equal(x1, x2) and np.array_equal(x1, x3)
    # inverse round trip
    winv = O.fr_inv(w)
    back = O.best_fft(x1, winv, log_n, 1)
    ninv = O.fr_from_int(pow(1 << log_n, -1, P.R_MOD))
    back = np.stack([O.fr_mul(x, ninv) for x in back])
    assert np.array_equal(back, a)


@pytest.mark.parametrize("k", [3, 5])
def test_domain_transforms_vs_model(k):
    d = O.EvaluationDomain(5, k)
    assert d.extended_k == k + 2 and d.quotient_poly_degree == 4
    n = 1 << k
    a = [rand_fr() for _ in range(n)]
    A = O.frs_from_ints(a)
    coeff = O.frs_to_ints(d.lagrange_to_coeff(A))
    assert coeff == P.lagrange_to_coeff(a, k)
    ext = O.frs_to_ints(d.coeff_to_extended(O.frs_from_ints(coeff)))
    assert ext == P.coeff_to_extended(coeff, k, k + 2)
    # coset evaluation: ext[i] = p(zeta * w_ext^i)
    wext = P.omega_for(k + 2)
    for i in (0, 1, 5, (1 << (k + 2)) - 1):
        assert ext[i] == P.eval_poly(coeff, P.ZETA * pow(wext, i, P.R_MOD) % P.R_MOD)
    back = O.frs_to_ints(d.extended_to_coeff(O.frs_from_ints(ext)))
    assert back == P.extended_to_coeff(ext, k + 2)
    assert back[:n] == coeff and all(v == 0 for v in back[n:])
    # t_evaluations: 1 / ((zeta w_ext^i)^n - 1)
    for i, t in enumerate(O.frs_to_ints(d.t_evaluations)):
        x = P.ZETA * pow(wext, i, P.R_MOD) % P.R_MOD
        assert t == pow(pow(x, n, P.R_MOD) - 1, -1, P.R_MOD)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 31, 32, 33, 100])
def test_multiexp_vs_double_and_add(n):
    bases_py = [P.g1_mul(P.G1_GEN, rng.randrange(1, 1 << 64)) for _ in range(n)]
    scal = [rand_fr() for _ in range(n)]
    if n >= 4:
        scal[0] = 0
        scal[1] = 1
        scal[2] = P.R_MOD - 1
        bases_py[3] = bases_py[2]  # duplicate base
    if n >= 33:
        bases_py[7] = None  # identity base
        scal[9] = scal[8]
    bases = np.stack([py_to_aff(p) for p in bases_py])
    S = O.frs_from_ints(scal)
    exp = P.msm(scal, bases_py)
    assert jac_to_py(O.multiexp_serial(S, bases)) == exp
    for t in (1, 3, 8):
        assert jac_to_py(O.best_multiexp(S, bases, t)) == exp


def test_best_multiexp_length_mismatch_asserts():
    with pytest.raises(AssertionError):
        O.best_multiexp(O.fill_fr(3, 1), O.fill_points(2, 1, 1))


def test_commit_is_evaluation_at_tau_and_lagrange_basis():
    k = 5
    n = 1 << k
    tau = rand_fr()
    g, gl = O.params_setup(k, O.fr_from_int(tau), threads=4)
    for i in (0, 1, n - 1):
        assert aff_to_py(g[i]) == P.g1_mul(P.G1_GEN, pow(tau, i, P.R_MOD))
    coeffs = [rand_fr() for _ in range(n)]
    C = O.commit(g, O.frs_from_ints(coeffs), threads=2)
    assert jac_to_py(C) == P.g1_mul(P.G1_GEN, P.eval_poly(coeffs, tau))
    # commit_lagrange(evals) == commit(coeffs) when evals = NTT(coeffs)
    evals = P.dft(coeffs, P.omega_for(k))
    CL = O.commit(gl, O.frs_from_ints(evals), threads=2)
    assert jac_to_py(CL) == jac_to_py(C)
    # Params::downsize path: g_to_lagrange(g) == g_lagrange
    assert np.array_equal(O.g_to_lagrange(g, k, threads=4), gl)


def test_eval_polynomial_and_kate_division():
    n = 37
    a = [rand_fr() for _ in range(n)]
    x, b = rand_fr(), rand_fr()
    assert O.fr_to_int(O.eval_polynomial(O.frs_from_ints(a), O.fr_from_int(x))) == P.eval_poly(a, x)
    q = O.frs_to_ints(O.kate_division(O.frs_from_ints(a), O.fr_from_int(b)))
    # a(X) - a(b) = q(X) (X - b)
    rem = P.eval_poly(a, b)
    prod = [0] * n
    for i, c in enumerate(q):
        prod[i + 1] = (prod[i + 1] + c) % P.R_MOD
        prod[i] = (prod[i] - b * c) % P.R_MOD
    prod[0] = (prod[0] + rem) % P.R_MOD
    assert prod == a
    ip = O.fr_to_int(O.compute_inner_product(O.frs_from_ints(a), O.frs_from_ints(a[::-1])))
    assert ip == sum(u * v for u, v in zip(a, a[::-1])) % P.R_MOD


def test_fill_generators_are_deterministic_and_valid():
    a = O.fill_fr(64, 0x5EEDB2000001)
    b = O.fill_fr(64, 0x5EEDB2000001)
    assert np.array_equal(a, b)
    assert all(O.limbs_to_int(x) < P.R_MOD for x in a)
    w = O.fill_fr(2000, 3, witness_like=True)
    zeros = sum(1 for x in w if not x.any())
    assert 1000 < zeros < 1400
    pts = O.fill_points(16, 5, 4)
    assert all(O.g1_affine_is_on_curve(p) for p in pts)
    assert np.array_equal(pts, O.fill_points(16, 5, 1))
