"""A miniature KZG-PLONK round trip over the hot path, verified by an INDEPENDENT pairing (tests/pairing_model.py).

One gate a*b - c = 0 on a 2^k-row domain, proved the way plonk::create_proof strings the path together
(/root/reference/integration/src/prove.rs:37-39 -> halo2_proofs create_proof):
  commit_lagrange(columns) -> lagrange_to_coeff -> coeff_to_extended -> gate values on the extended coset (GraphEvaluator)
  -> divide by X^n - 1 -> extended_to_coeff -> commit(h) -> evaluate everything at a challenge x -> kate_division openings
and then checked the way a verifier would: every opening satisfies  e(C - y*G, G2) = e(W, [tau - x]G2)  and the gate
identity a(x)*b(x) - c(x) = h(x)*(x^n - 1) holds.  A wrong MSM, NTT, coset extension, division or evaluation anywhere on the
path breaks a pairing equation; neither the oracle nor the product is asked to vouch for itself.

Runs twice: on the CPU with the oracle's restatement (pins the oracle), on the GPU through the C ABI (pins the product).
"""
import random

import numpy as np
import pytest

from oracle import oracle as O
from pairing_model import G1_GEN, G2_GEN, Q, g1_add, g1_mul, g2_add, g2_mul, g2_neg, pairing_check
from quotient_programs import C_MUL, C_SUB, R_MOD, S_ADVICE, S_INTER, ZETA, omega_of

K, J = 7, 5
N = 1 << K
TAU = 0x1D7C3A9B5E2F4061_8899AABBCCDDEEFF_0123456789ABCDEF % R_MOD


def pt_from_jac(j):
    a = O.g1_to_affine(j)
    x, y = O.fq_to_int(a[:4]), O.fq_to_int(a[4:])
    return None if (x, y) == (0, 0) else (x, y)


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q)


def verify_opening(commitment, x, y, witness) -> bool:
    """e(C - y*G, G2) * e(-W, [tau]G2 - [x]G2) == 1"""
    lhs = g1_add(commitment, g1_neg(g1_mul(G1_GEN, y))) if y else commitment
    s_minus_x = g2_add(g2_mul(G2_GEN, TAU), g2_neg(g2_mul(G2_GEN, x)))
    return pairing_check([(lhs, G2_GEN), (g1_neg(witness), s_minus_x)])


def witness(seed):
    rng = random.Random(seed)
    a = [rng.randrange(R_MOD) for _ in range(N)]
    b = [rng.randrange(R_MOD) for _ in range(N)]
    c = [x * y % R_MOD for x, y in zip(a, b)]
    return a, b, c


def t_inv_column(ek):
    we = omega_of(ek)
    ext = 1 << (ek - K)
    vals = [pow((pow(ZETA, N, R_MOD) * pow(we, N * i, R_MOD) - 1) % R_MOD, -1, R_MOD) for i in range(ext)]
    return [vals[i % ext] for i in range(1 << ek)]


class OracleBackend:
    """the CPU restatement (oracle/) driven through the same steps"""

    def __init__(self):
        self.dom = O.EvaluationDomain(J, K)
        self.g, self.g_lagrange = O.params_setup(K, O.fr_from_int(TAU), threads=4)

    def commit_lagrange(self, col): return pt_from_jac(O.commit(self.g_lagrange, col))
    def commit(self, poly): return pt_from_jac(O.commit(self.g[: len(poly)], poly))
    def lagrange_to_coeff(self, col): return self.dom.lagrange_to_coeff(col)
    def coeff_to_extended(self, p): return self.dom.coeff_to_extended(p)

    def gate(self, a_ext, b_ext, c_ext):
        z = O.fr_from_int(0)
        e = np.zeros((0, 4), np.uint64)
        prog = [(C_MUL, (S_ADVICE, 0, 0), (S_ADVICE, 1, 0), None), (C_SUB, (S_INTER, 0, 0), (S_ADVICE, 2, 0), None)]
        return O.graph_evaluate(prog, e, [0], [], [a_ext, b_ext, c_ext], [], e, z, z, z, z, None, np.zeros_like(a_ext),
                                self.dom.extended_k, 1)

    def pointwise_mul(self, u, v): return np.stack([O.fr_mul(x, y) for x, y in zip(u, v)])
    def extended_to_coeff(self, v): return self.dom.extended_to_coeff(v)
    def eval(self, p, x): return O.fr_to_int(O.eval_polynomial(p, O.fr_from_int(x)))
    def kate(self, p, x): return O.kate_division(p, O.fr_from_int(x))


class DeviceBackend:
    """the product through the C ABI: SRS generated, transformed and used on the device"""

    def __init__(self, zk, ctx):
        self.zk, self.ctx = zk, ctx
        self.dom = zk.EvaluationDomain(ctx, J, K)
        powers = O.frs_from_ints([pow(TAU, i, R_MOD) for i in range(N)])
        g = ctx.g1_generator_mul_batch(powers)            # ParamsKZG::setup: g[i] = [tau^i] G
        gl = ctx.g_to_lagrange(g, K)                      # g_lagrange by the G1 inverse FFT
        self.params = zk.ParamsKZG(ctx, K, g, gl)

    def commit_lagrange(self, col): return pt_from_jac(self.params.commit_lagrange(col))
    def commit(self, poly): return pt_from_jac(self.params.commit(poly))
    def lagrange_to_coeff(self, col): return self.dom.lagrange_to_coeff(col.copy())
    def coeff_to_extended(self, p): return self.dom.coeff_to_extended(p)

    def gate(self, a_ext, b_ext, c_ext):
        import torch

        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()
        prog = [(C_MUL, (S_ADVICE, 0, 0), (S_ADVICE, 1, 0), None), (C_SUB, (S_INTER, 0, 0), (S_ADVICE, 2, 0), None)]
        g = self.ctx.graph(prog, np.zeros((0, 4), np.uint64), [0])
        cols = [dev(a_ext), dev(b_ext), dev(c_ext)]
        out = dev(np.zeros_like(a_ext))
        torch.cuda.synchronize()
        g.evaluate(out, self.dom.extended_k, 1, advice=cols)
        self.ctx.synchronize()
        return out.cpu().numpy().view(np.uint64)

    def pointwise_mul(self, u, v): return self.ctx.poly_mul(u, v)
    def extended_to_coeff(self, v): return self.dom.extended_to_coeff(v.copy())
    def eval(self, p, x): return O.fr_to_int(self.ctx.eval_polynomial(p, O.fr_from_int(x)))
    def kate(self, p, x): return self.ctx.kate_division(p, O.fr_from_int(x))


def run_round_trip(be, seed):
    a, b, c = witness(seed)
    cols = {n: O.frs_from_ints(v) for n, v in (("a", a), ("b", b), ("c", c))}
    commits = {n: be.commit_lagrange(v) for n, v in cols.items()}
    coeffs = {n: be.lagrange_to_coeff(v) for n, v in cols.items()}
    for n in cols:  # commit(coefficients) and commit_lagrange(values) are the same group element
        assert be.commit(coeffs[n]) == commits[n]
    ext = {n: be.coeff_to_extended(coeffs[n]) for n in cols}
    ek = K + 2
    num = be.gate(ext["a"], ext["b"], ext["c"])
    quot = be.pointwise_mul(num, O.frs_from_ints(t_inv_column(ek)))
    h_all = np.asarray(be.extended_to_coeff(quot))
    assert not h_all[N - 1:].any(), "a*b - c must be divisible by X^n - 1: deg h <= n - 2"
    coeffs["h"] = np.ascontiguousarray(h_all[:N])
    commits["h"] = be.commit(coeffs["h"])
    x = random.Random(seed + 1).randrange(R_MOD)
    evals = {n: be.eval(coeffs[n], x) for n in coeffs}
    assert (evals["a"] * evals["b"] - evals["c"]) % R_MOD == evals["h"] * (pow(x, N, R_MOD) - 1) % R_MOD
    for n in coeffs:
        w = be.commit(be.kate(coeffs[n], x))
        assert verify_opening(commits[n], x, evals[n], w), n
    # and the verifier is not vacuous: a shifted evaluation is rejected
    w = be.commit(be.kate(coeffs["a"], x))
    assert not verify_opening(commits["a"], x, (evals["a"] + 1) % R_MOD, w)
    return commits, evals


def test_round_trip_with_the_oracle_verifies_under_the_pairing():
    run_round_trip(OracleBackend(), 11)


@pytest.mark.gpu
def test_round_trip_on_the_device_verifies_under_the_pairing(zk, ctx):
    commits_d, evals_d = run_round_trip(DeviceBackend(zk, ctx), 11)
    commits_o, evals_o = run_round_trip(OracleBackend(), 11)
    assert commits_d == commits_o and evals_d == evals_o  # and both agree, point for point
