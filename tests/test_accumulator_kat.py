"""Verifier-side known-answer test on the reference's own artefacts.

The KZG accumulators the reference ships -- the first 384 B of release-v0.13.1/proof.data (12 limbs of 88 bits,
/root/reference/integration/tests/unit_tests.rs:32) and the accumulator instances of
integration/tests/test_data/full_proof_1.json -- must satisfy the pairing equation of
release-v0.13.1/evm_verifier.yul:1230-1240:   e(lhs, G2) * e(rhs, X) == 1,  X = the G2 constant at 0x50c0 (= -[s]G2).
That equation is linear in (lhs, rhs), so any multi-scalar multiplication over pairs derived from those accumulators must
produce a pair that satisfies it again (this is how snark-verifier folds accumulators).  That gives the one cryptographic
check the reference's artefacts allow on MSM *results*: a wrong bucket, window or carry in the MSM breaks the pairing
equation with overwhelming probability.  The pairing itself is computed by an independent big-integer model
(tests/pairing_model.py); nothing here trusts the oracle's or the product's own arithmetic to check itself.
"""
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle as O
from pairing_model import (F12, G1_GEN, G2_GEN, Q, R, g1_add, g1_mul, g2_mul, g2_neg, g2_on_curve, pairing, pairing_check)

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_fixtures.json")))


def yul_g2_constants():
    """evm_verifier.yul:1230-1239: two G2 points in EIP-197 word order (x_c1, x_c0, y_c1, y_c0)"""
    vals = [int(m.group(1), 16) for l in GOLD["evm_verifier_yul"]["lines1230_1240"] for m in [re.search(r", (0x[0-9a-f]{64})\)", l)] if m]
    assert len(vals) == 8
    g2 = ((vals[1], vals[0]), (vals[3], vals[2]))
    x = ((vals[5], vals[4]), (vals[7], vals[6]))
    return g2, x


def accumulator(hx: str):
    b = bytes.fromhex(hx)
    w = [int.from_bytes(b[32 * i: 32 * i + 32], "big") for i in range(12)]
    c = [w[3 * i] + (w[3 * i + 1] << 88) + (w[3 * i + 2] << 176) for i in range(4)]
    return (c[0], c[1]), (c[2], c[3])


ACCS = [accumulator(GOLD["files"]["proof.data"]["accumulator_hex"]), accumulator(GOLD["full_proof_1"]["instances_accumulator_hex"])]


def to_affine_limbs(p):
    return np.concatenate([O.fq_from_int(p[0]), O.fq_from_int(p[1])])


def from_jacobian(j):
    a = O.g1_to_affine(j)
    return O.fq_to_int(a[:4]), O.fq_to_int(a[4:])


def test_pairing_model_is_bilinear_and_non_degenerate():
    e1 = pairing(G2_GEN, G1_GEN)
    assert not e1 == F12.one() and e1 ** R == F12.one()
    assert pairing(G2_GEN, g1_mul(G1_GEN, 5)) == e1 ** 5
    assert pairing(g2_mul(G2_GEN, 7), G1_GEN) == e1 ** 7
    assert pairing_check([(g1_mul(G1_GEN, 6), G2_GEN), (G1_GEN, g2_neg(g2_mul(G2_GEN, 6)))])
    assert not pairing_check([(g1_mul(G1_GEN, 6), G2_GEN), (G1_GEN, g2_neg(g2_mul(G2_GEN, 5)))])


def test_reference_accumulators_satisfy_the_evm_verifier_pairing_equation():
    g2, x = yul_g2_constants()
    assert g2 == G2_GEN and g2_on_curve(x)
    for lhs, rhs in ACCS:
        assert pairing_check([(lhs, g2), (rhs, x)])
        assert not pairing_check([(lhs, g2), (rhs, g2_neg(x))])          # the constant is -[s]G2, not [s]G2
        assert not pairing_check([(g1_add(lhs, G1_GEN), g2), (rhs, x)])  # and the check is not vacuous


def kat_bases(n: int):
    """n base pairs (L_i, R_i) = t_i * (lhs_j, rhs_j), j = i mod 2, small t_i: every pair satisfies the equation"""
    bl, br = [], []
    for i in range(n):
        lhs, rhs = ACCS[i % 2]
        t = 1 + (i * 2654435761) % 1009
        bl.append(g1_mul(lhs, t))
        br.append(g1_mul(rhs, t))
    return np.stack([to_affine_limbs(p) for p in bl]), np.stack([to_affine_limbs(p) for p in br])


def test_oracle_msm_results_satisfy_the_reference_pairing_equation():
    """best_multiexp (the oracle's restatement, serial and chunked) over accumulator-derived bases folds to a valid pair."""
    g2, x = yul_g2_constants()
    n = 96
    bl, br = kat_bases(n)
    coeffs = O.fill_fr(n, 0xACC)
    for threads in (1, 4):
        lo = from_jacobian(O.best_multiexp(coeffs, bl, threads=threads))
        ro = from_jacobian(O.best_multiexp(coeffs, br, threads=threads))
        assert pairing_check([(lo, g2), (ro, x)])
    # a single wrong coefficient on one side breaks it
    bad = coeffs.copy()
    bad[5] = O.fr_from_int(O.fr_to_int(bad[5]) + 1)
    ro_bad = from_jacobian(O.best_multiexp(bad, br))
    assert not pairing_check([(lo, g2), (ro_bad, x)])


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 96, 1 << 12, 1 << 16])
def test_device_msm_results_satisfy_the_reference_pairing_equation(ctx, n):
    """The CUDA Pippenger (plain and precomputed-table paths) over accumulator-derived bases, same scalars on both sides:
    the two results must again satisfy e(L, G2) * e(R, X) == 1 with the reference's X."""
    g2, x = yul_g2_constants()
    m = min(n, 128)  # distinct base pairs (host big-int scalar multiplications); larger n cycle through them
    bl, br = kat_bases(m)
    idx = np.arange(n) % m
    bl, br = bl[idx], br[idx]
    coeffs = O.fill_fr(n, 0xACC + n)
    lo = from_jacobian(ctx.best_multiexp(coeffs, bl))
    ro = from_jacobian(ctx.best_multiexp(coeffs, br))
    assert pairing_check([(lo, g2), (ro, x)])
    sl, sr = ctx.srs_register(bl), ctx.srs_register(br)  # resident SRS handles (precomputed 2^(c*w) tables at n = 2^16)
    lo2, ro2 = from_jacobian(sl.msm(coeffs)), from_jacobian(sr.msm(coeffs))
    assert (lo2, ro2) == (lo, ro)
    sl.release()
    sr.release()
    bad = coeffs.copy()
    bad[n - 1] = O.fr_from_int(O.fr_to_int(bad[n - 1]) + 1)
    assert not pairing_check([(lo, g2), (from_jacobian(ctx.best_multiexp(bad, br)), x)])
