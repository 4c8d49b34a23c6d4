"""The oracle's quotient-construction restatement against the committed vectors of tests/golden/quotient_vectors.json
(written by tests/golden/make_quotient_vectors.py, cross-checked against the big-integer model at generation time), and the
product's host-emulated lowering + interpreter against the same bytes.  CPU only."""
import json
import os

import numpy as np

from h_terms_programs import permutation_terms_program
from oracle import oracle as O
from test_graph_host_emul import host_eval, lib  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "quotient_vectors.json")))


def un(h):
    return np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, 4).copy()


def un1(h):
    return un(h)[0]


def calcs_of(g):
    return [(op, tuple(a), tuple(b) if b else None, [tuple(p) for p in ps] if ps else None) for op, a, b, ps in g["calcs"]]


def test_graph_vector(lib):
    g = GOLD["graph"]
    args = (calcs_of(g), un(g["constants"]), g["rotations"], [un(c) for c in g["fixed"]], [un(c) for c in g["advice"]],
            [un(c) for c in g["instance"]], un(g["challenges"]))
    got = O.graph_evaluate(*args, un1(g["beta"]), un1(g["gamma"]), un1(g["theta"]), un1(g["y"]), un1(g["extended_omega"]), un(g["previous"]),
                           g["log_size"], g["rot_scale"])
    assert np.array_equal(got, un(g["output"]))
    rc, he, _, err = host_eval(lib, *args, [un1(g["beta"]), un1(g["gamma"]), un1(g["theta"]), un1(g["y"])], un1(g["extended_omega"]),
                               un(g["previous"]), g["log_size"], g["rot_scale"])
    assert rc == 0, err
    assert np.array_equal(he, un(g["output"]))


def test_permutation_product_vector():
    g = GOLD["permutation_product"]
    z = O.permutation_product([un(c) for c in g["values"]], [un(c) for c in g["sigma"]], un1(g["beta"]), un1(g["gamma"]),
                              un1(g["delta_omega_start"]), un1(g["delta"]), un1(g["omega"]), g["k"], un1(g["z_init"]))
    assert np.array_equal(z, un(g["z"]))


def test_logup_vector():
    g = GOLD["logup"]
    phi = O.logup_running_sum([un(c) for c in g["inputs"]], un(g["table"]), un(g["m"]), un1(g["beta"]), g["k"], un1(g["phi_init"]))
    assert np.array_equal(phi, un(g["phi"]))


def test_permutation_h_terms_vector(lib):
    g = GOLD["permutation_h_terms"]
    ek, ext = g["extended_k"], g["extended_k"] - g["k"]
    z, v, s = [un(c) for c in g["z"]], [un(c) for c in g["values"]], [un(c) for c in g["sigma"]]
    l0, ll, la = un(g["l0"]), un(g["l_last"]), un(g["l_active_row"])
    want = O.permutation_h_terms(z, g["chunk_len"], v, s, l0, ll, la, un1(g["beta"]), un1(g["gamma"]), un1(g["y"]), un1(g["delta"]),
                                 un1(g["extended_omega"]), g["last_rotation"], un(g["previous"]), ek, 1 << ext)
    assert np.array_equal(want, un(g["output"]))
    calcs, constants, rotations = permutation_terms_program(len(z), g["chunk_len"], len(v), g["last_rotation"])
    zero = O.fr_from_int(0)
    rc, he, _, err = host_eval(lib, calcs, O.frs_from_ints(constants), rotations, s + [l0, ll, la], z + v, [], np.zeros((0, 4), np.uint64),
                               [un1(g["beta"]), un1(g["gamma"]), zero, un1(g["y"])], un1(g["extended_omega"]), un(g["previous"]), ek, 1 << ext)
    assert rc == 0, err
    assert np.array_equal(he, un(g["output"]))
