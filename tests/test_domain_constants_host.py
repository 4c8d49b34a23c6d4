"""The product's host-side EvaluationDomain constants (scroll-prover_b200/__init__.py, computed without touching the device)
against the reference's own dump of EvaluationDomain::new(5, 25) in release-v0.13.1/chunk.protocol and against the oracle for
every degree the reference uses (20, 21, 24, 25, 26: params-sha256sum:1-5, integration/configs/layer*.config:3).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_fixtures.json")))


def test_k25_constants_equal_the_reference_dump(zk):
    dom = zk.EvaluationDomain(None, 5, 25)  # the constructor only computes constants
    ref = GOLD["chunk_protocol"]["domain"]
    assert ref["k"] == 25 and ref["n"] == 1 << 25
    assert np.array_equal(dom.omega, np.array(ref["gen"], dtype=np.uint64))
    assert np.array_equal(dom.omega_inv, np.array(ref["gen_inv"], dtype=np.uint64))
    assert np.array_equal(dom.ifft_divisor, np.array(ref["n_inv"], dtype=np.uint64))
    assert dom.extended_k == 27 and dom.quotient_poly_degree == GOLD["chunk_protocol"]["quotient_num_chunk"]


@pytest.mark.parametrize("k", [3, 10, 20, 21, 24, 25, 26])
def test_constants_equal_the_oracles_domain(zk, k):
    dom, od = zk.EvaluationDomain(None, 5, k), O.EvaluationDomain(5, k)
    s = od._s
    assert dom.extended_k == od.extended_k == k + 2
    for mine, theirs in ((dom.omega, s.omega), (dom.omega_inv, s.omega_inv), (dom.extended_omega, s.extended_omega),
                         (dom.extended_omega_inv, s.extended_omega_inv), (dom.g_coset, s.g_coset), (dom.g_coset_inv, s.g_coset_inv),
                         (dom.ifft_divisor, s.ifft_divisor), (dom.extended_ifft_divisor, s.extended_ifft_divisor)):
        assert np.array_equal(np.asarray(mine, dtype=np.uint64), np.array(list(theirs), dtype=np.uint64))


def test_field_element_helpers_round_trip(zk):
    for v in (0, 1, 2, zk.R_MOD - 1, 0xDEADBEEF << 200):
        assert zk.fr_to_int(zk.fr_from_int(v)) == v % zk.R_MOD
        assert np.array_equal(zk.fr_from_int(v), O.fr_from_int(v))
