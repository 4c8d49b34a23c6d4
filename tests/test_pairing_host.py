"""The product's host-side pairing (scroll-prover_b200/pairing_bn254.hpp) through an EIP-197-shaped C shim:
bilinearity, agreement with the independent big-integer model on accept/reject decisions, and the reference's own
accumulators (release-v0.13.1/proof.data, full_proof_1.json) against the G2 constants of evm_verifier.yul:1230-1239.
CPU only: the header runs on the host build of csrc/ff.cuh."""
import ctypes as C
import json
import os
import random
import subprocess

import pytest

from pairing_model import G1_GEN, G2_GEN, Q, R, g1_add, g1_mul, g2_add, g2_mul, g2_neg, pairing_check
from test_accumulator_kat import ACCS, GOLD, yul_g2_constants

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "pairing_host.cpp")
SO = os.path.join(HERE, "host_emul", "libpairing_host.so")
HDRS = [os.path.join(HERE, "..", "scroll-prover_b200", p) for p in ("pairing_bn254.hpp", os.path.join("csrc", "ff.cuh"))]


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(p) for p in [SRC] + HDRS):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return C.CDLL(SO)


def be32(v):
    return int(v).to_bytes(32, "big")


def enc_g1(p):
    return be32(0) * 2 if p is None else be32(p[0]) + be32(p[1])


def enc_g2(q):
    if q is None:
        return be32(0) * 4
    (x0, x1), (y0, y1) = q
    return be32(x1) + be32(x0) + be32(y1) + be32(y0)  # EIP-197 order: imaginary part first


def check(lib, pairs):
    buf = b"".join(enc_g1(p) + enc_g2(q) for p, q in pairs)
    return lib.pairing_host_eip197(buf, len(pairs))


def test_bilinearity_and_non_degeneracy(lib):
    a, b = 0x1234567, 0xABCDEF01
    # e(aP, Q) * e(P, -aQ) == 1 ; e(aP, bQ) == e(abP, Q) ; e(P, Q) != 1
    assert check(lib, [(g1_mul(G1_GEN, a), G2_GEN), (G1_GEN, g2_neg(g2_mul(G2_GEN, a)))]) == 1
    assert check(lib, [(g1_mul(G1_GEN, a), g2_mul(G2_GEN, b)), (g1_mul(G1_GEN, a * b % R), g2_neg(G2_GEN))]) == 1
    assert check(lib, [(G1_GEN, G2_GEN)]) == 0
    assert check(lib, [(g1_mul(G1_GEN, a), G2_GEN), (G1_GEN, g2_neg(g2_mul(G2_GEN, a + 1)))]) == 0
    pa = enc_g1(g1_mul(G1_GEN, 6)) + enc_g2(G2_GEN)
    pb = enc_g1(g1_mul(G1_GEN, 2)) + enc_g2(g2_mul(G2_GEN, 3))
    pc = enc_g1(g1_mul(G1_GEN, 2)) + enc_g2(g2_mul(G2_GEN, 4))
    assert lib.pairing_host_equal(pa, pb) == 1 and lib.pairing_host_equal(pa, pc) == 0


def test_identities_and_degenerate_pairs(lib):
    assert check(lib, []) == 1
    assert check(lib, [(None, G2_GEN)]) == 1 and check(lib, [(G1_GEN, None)]) == 1
    p = g1_mul(G1_GEN, 77)
    # P and -P against the same Q cancel; the doubling and the vertical-line branches of the Miller loop are both taken
    assert check(lib, [(p, G2_GEN), ((p[0], (-p[1]) % Q), G2_GEN)]) == 1
    assert check(lib, [(p, g2_mul(G2_GEN, 2)), (g1_mul(p, 2), g2_neg(G2_GEN))]) == 1


def test_malformed_points_are_rejected(lib):
    bad_g1 = (1, 3)  # not on y^2 = x^3 + 3
    assert check(lib, [(bad_g1, G2_GEN)]) == -1
    assert check(lib, [((Q, 2), G2_GEN)]) == -1  # coordinate not reduced
    (x0, x1), y = G2_GEN
    assert check(lib, [(G1_GEN, ((x0 + 1, x1), y))]) == -1


def test_decisions_agree_with_the_bigint_model(lib):
    rng = random.Random(99)
    for trial in range(6):
        a, b, c = (rng.randrange(1, R) for _ in range(3))
        good = trial % 2 == 0
        d = a * b % R if good else (a * b + 1) % R
        # e(aG, bH) * e(cG, H) * e(-(d + c)G, H) == 1  iff  d == ab
        pairs = [(g1_mul(G1_GEN, a), g2_mul(G2_GEN, b)), (g1_mul(G1_GEN, c), G2_GEN), (g1_mul(G1_GEN, (R - (d + c) % R) % R), G2_GEN)]
        assert check(lib, pairs) == (1 if good else 0)
        assert pairing_check(pairs) == good


def test_reference_accumulators_are_accepted_by_the_host_verifier(lib):
    g2, x = yul_g2_constants()
    for hx, (lhs, rhs) in zip((GOLD["files"]["proof.data"]["accumulator_hex"], GOLD["full_proof_1"]["instances_accumulator_hex"]), ACCS):
        raw = bytes.fromhex(hx)
        assert lib.pairing_host_accumulator(raw, enc_g2(g2), enc_g2(x)) == 1            # as shipped: 12 limbs of 88 bits
        assert lib.pairing_host_accumulator(raw, enc_g2(g2), enc_g2(g2_neg(x))) == 0    # the constant is -[s]G2
        assert check(lib, [(lhs, g2), (rhs, x)]) == 1                                    # same through the EIP-197 entry
        tampered = bytearray(raw)
        tampered[31] ^= 1                                                                # lowest limb of lhs.x
        assert lib.pairing_host_accumulator(bytes(tampered), enc_g2(g2), enc_g2(x)) in (0, -1)
    # a limb word with bits above 2^88 is malformed
    bad = bytearray(bytes.fromhex(GOLD["files"]["proof.data"]["accumulator_hex"]))
    bad[0] = 1
    assert lib.pairing_host_accumulator(bytes(bad), enc_g2(g2), enc_g2(x)) == -1


def test_folded_accumulators_stay_valid(lib):
    """snark-verifier's accumulation: a random linear combination of valid accumulators is a valid accumulator"""
    g2, x = yul_g2_constants()
    r1, r2 = 0xDEADBEEF12345, 0xFEEDFACE6789
    (l1, h1), (l2, h2) = ACCS
    lhs = g1_add(g1_mul(l1, r1), g1_mul(l2, r2))
    rhs = g1_add(g1_mul(h1, r1), g1_mul(h2, r2))
    assert check(lib, [(lhs, g2), (rhs, x)]) == 1
    assert check(lib, [(lhs, g2), (g1_add(rhs, G1_GEN), x)]) == 0
