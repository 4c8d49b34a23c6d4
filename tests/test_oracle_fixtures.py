"""Pins the CPU oracle against every fixture the reference ships for this path.

The fixtures were extracted from /root/reference by tests/golden/make_golden.py; nothing here
reads /root/reference at run time.
"""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_fixtures.json")))


def test_moduli_match_evm_verifier_yul():
    # release-v0.13.1/evm_verifier.yul:17-18
    f_p = int(GOLD["evm_verifier_yul"]["line17"].split(":=")[1].strip(), 16)
    f_q = int(GOLD["evm_verifier_yul"]["line18"].split(":=")[1].strip(), 16)
    assert f_p == O.Q_MOD == P.Q_MOD
    assert f_q == O.R_MOD == P.R_MOD
    # the C constants: 0 - 1 == p - 1 in both fields
    zero = np.zeros(4, np.uint64)
    assert int.from_bytes(O.fr_to_repr(O.fr_sub(zero, O.const_fr("fr_ONE"))), "little") == f_q - 1
    assert int.from_bytes(O.fq_to_repr(O.fq_sub(zero, O.const_fr("fq_ONE"))), "little") == f_p - 1


def test_montgomery_constants():
    assert O.limbs_to_int(O.const_fr("fr_ONE")) == (1 << 256) % O.R_MOD
    assert O.limbs_to_int(O.const_fr("fr_R2")) == (1 << 512) % O.R_MOD
    assert O.limbs_to_int(O.const_fr("fq_ONE")) == (1 << 256) % O.Q_MOD
    assert O.limbs_to_int(O.const_fr("fq_R2")) == (1 << 512) % O.Q_MOD
    assert O.fr_to_int(O.const_fr("fr_ROOT_OF_UNITY")) == P.ROOT_OF_UNITY
    assert O.fr_to_int(O.const_fr("fr_ZETA")) == P.ZETA
    assert pow(P.ZETA, 3, P.R_MOD) == 1 and P.ZETA != 1
    assert pow(P.ROOT_OF_UNITY, 1 << 28, P.R_MOD) == 1 and pow(P.ROOT_OF_UNITY, 1 << 27, P.R_MOD) != 1


def test_domain_matches_chunk_protocol():
    # release-v0.13.1/chunk.protocol "domain": raw Montgomery limbs of n_inv, gen, gen_inv at k = 25
    dom = GOLD["chunk_protocol"]["domain"]
    assert dom["k"] == 25 and dom["n"] == 1 << 25
    assert GOLD["chunk_protocol"]["quotient_num_chunk"] == 4
    d = O.EvaluationDomain(5, 25)  # j = num_chunk + 1
    assert d.extended_k == 27
    assert list(map(int, d.omega)) == dom["gen"]
    assert list(map(int, d.omega_inv)) == dom["gen_inv"]
    assert list(map(int, d.ifft_divisor)) == dom["n_inv"]
    # omega(k=25) = ROOT_OF_UNITY^8, extended_omega^4 = omega
    assert O.fr_to_int(d.omega) == pow(P.ROOT_OF_UNITY, 8, P.R_MOD)
    assert pow(O.fr_to_int(d.extended_omega), 4, P.R_MOD) == O.fr_to_int(d.omega)


def _vk_points(hexstr):
    b = bytes.fromhex(hexstr)
    k = int.from_bytes(b[0:4], "big")
    nfixed = int.from_bytes(b[4:8], "big")
    pts = [b[8 + 32 * i: 40 + 32 * i] for i in range((len(b) - 8) // 32)]
    return k, nfixed, pts


def test_preprocessed_points_on_curve_and_match_vk_chunk():
    pre = GOLD["chunk_protocol"]["preprocessed"]
    k, nfixed, pts = _vk_points(GOLD["files"]["vk_chunk.vkey"]["hex"])
    assert k == 25 and nfixed == 4 and len(pts) == 7 == len(pre)
    for p, comp in zip(pre, pts):
        aff = np.array(p["x"] + p["y"], dtype=np.uint64)
        assert O.g1_affine_is_on_curve(aff)
        assert P.on_curve((O.fq_to_int(aff[:4]), O.fq_to_int(aff[4:])))
        assert O.g1_compress(aff) == comp
        assert P.compress((O.fq_to_int(aff[:4]), O.fq_to_int(aff[4:]))) == comp
        back = O.g1_decompress(comp)
        assert back is not None and np.array_equal(back, aff)


@pytest.mark.parametrize("name,k,npts", [("vk_chunk.vkey", 25, 7), ("vk_batch.vkey", 26, 9), ("vk_bundle.vkey", 26, 7)])
def test_vk_files_decode(name, k, npts):
    kk, nfixed, pts = _vk_points(GOLD["files"][name]["hex"])
    assert kk == k and nfixed == 4 and len(pts) == npts
    for comp in pts:
        aff = O.g1_decompress(comp)
        assert aff is not None and O.g1_affine_is_on_curve(aff)
        assert O.g1_compress(aff) == comp


def _acc_points(b384: bytes):
    words = [int.from_bytes(b384[32 * i: 32 * i + 32], "big") for i in range(12)]
    coords = [words[3 * i] + (words[3 * i + 1] << 88) + (words[3 * i + 2] << 176) for i in range(4)]
    return (coords[0], coords[1]), (coords[2], coords[3])


def test_kzg_accumulators_on_curve():
    # proof.data first 384 B (integration/tests/unit_tests.rs:32) and full_proof_1 instances
    for hx in (GOLD["files"]["proof.data"]["accumulator_hex"], GOLD["full_proof_1"]["instances_accumulator_hex"]):
        lhs, rhs = _acc_points(bytes.fromhex(hx))
        for pt in (lhs, rhs):
            assert P.on_curve(pt)
            aff = np.concatenate([O.fq_from_int(pt[0]), O.fq_from_int(pt[1])])
            assert O.g1_affine_is_on_curve(aff)


def test_chunk_proof_commitments_decode():
    # layer-2 proof: 5 witness + 4 quotient commitments, 17 evals, 2 SHPLONK points
    fp = GOLD["full_proof_1"]
    assert fp["proof_len"] == 896
    proof = bytes.fromhex(fp["proof_hex"])
    assert sum(GOLD["chunk_protocol"]["num_witness"]) == 5
    point_words = list(range(0, 9)) + [26, 27]
    for w in point_words:
        aff = O.g1_decompress(proof[32 * w: 32 * w + 32])
        assert aff is not None and O.g1_affine_is_on_curve(aff), f"word {w}"
    for w in range(9, 26):  # evaluations are canonical Fr
        assert int.from_bytes(proof[32 * w: 32 * w + 32], "little") < P.R_MOD
    assert fp["vk_hex"] == GOLD["files"]["vk_chunk.vkey"]["hex"]


def test_layer_configs_define_sizes():
    ks = [GOLD["layer_configs"][f"layer{i}"]["degree"] for i in range(1, 7)]
    assert ks == [24, 25, 21, 26, 21, 26]
