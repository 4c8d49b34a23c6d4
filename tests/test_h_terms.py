"""evaluate_h's permutation and lookup sections as GraphEvaluator programs (tests/h_terms_programs.py) against the oracle's
restatement of the upstream loops (oracle/halo2_quotient.c: halo2_permutation_h_terms, halo2_logup_h_terms):
on the CPU through the host-emulated lowering + interpreter, on the GPU through the C ABI."""
import random

import numpy as np
import pytest

from h_terms_programs import logup_terms_program, permutation_terms_program
from oracle import oracle as O
from quotient_programs import DELTA, R_MOD, omega_of
from test_graph_host_emul import host_eval, lib  # noqa: F401  (lib is the host-emulation fixture)


def perm_case(seed, k, ext, n_sets, chunk_len, n_cols, blinding):
    ek = k + ext
    size = 1 << ek
    rng = random.Random(seed)
    col = lambda: O.fill_fr(size, rng.randrange(1 << 30))
    z = [col() for _ in range(n_sets)]
    v = [col() for _ in range(n_cols)]
    s = [col() for _ in range(n_cols)]
    l0, l_last, l_act = col(), col(), col()
    beta, gamma, y = O.fill_fr(3, seed + 5)
    prev = col()
    last_rotation = -(blinding + 1)
    we = O.fr_from_int(omega_of(ek))
    want = O.permutation_h_terms(z, chunk_len, v, s, l0, l_last, l_act, beta, gamma, y, O.fr_from_int(DELTA), we, last_rotation, prev,
                                 ek, 1 << ext)
    calcs, constants, rotations = permutation_terms_program(n_sets, chunk_len, n_cols, last_rotation)
    return dict(calcs=calcs, constants=O.frs_from_ints(constants), rotations=rotations, fixed=s + [l0, l_last, l_act], advice=z + v,
                beta=beta, gamma=gamma, y=y, we=we, prev=prev, ek=ek, rot_scale=1 << ext, want=want)


def logup_case(seed, k, ext, n_inputs):
    ek = k + ext
    size = 1 << ek
    rng = random.Random(seed)
    col = lambda: O.fill_fr(size, rng.randrange(1 << 30))
    f = [col() for _ in range(n_inputs)]
    table, m, phi = col(), col(), col()
    l0, l_last, l_act = col(), col(), col()
    beta, y = O.fill_fr(2, seed + 9)
    prev = col()
    want = O.logup_h_terms(f, table, m, phi, l0, l_last, l_act, beta, y, prev, ek, 1 << ext)
    calcs, constants, rotations = logup_terms_program(n_inputs)
    return dict(calcs=calcs, constants=O.frs_from_ints(constants), rotations=rotations, fixed=[l0, l_last, l_act], advice=f + [table, m, phi],
                beta=beta, gamma=O.fr_from_int(0), y=y, we=None, prev=prev, ek=ek, rot_scale=1 << ext, want=want)


CASES = [("perm", (1, 3, 2, 1, 3, 2, 5)), ("perm", (2, 4, 2, 3, 3, 8, 5)), ("perm", (3, 3, 1, 2, 2, 3, 0)), ("perm", (4, 2, 2, 4, 1, 4, 1)),
         ("logup", (5, 4, 2, 1)), ("logup", (6, 3, 2, 3)), ("logup", (7, 3, 1, 6))]


def build(kind, args):
    return perm_case(*args) if kind == "perm" else logup_case(*args)


@pytest.mark.parametrize("kind,args", CASES)
def test_h_term_programs_match_the_upstream_loops_host_emulated(lib, kind, args):
    c = build(kind, args)
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    rc, got, (n_instr, n_slots), err = host_eval(lib, c["calcs"], c["constants"], c["rotations"], c["fixed"], c["advice"], [], e,
                                                 [c["beta"], c["gamma"], z, c["y"]], c["we"], c["prev"], c["ek"], c["rot_scale"])
    assert rc == 0, err
    assert np.array_equal(got, c["want"])
    # the y-fold keeps one term live at a time whatever the number of sets; a lookup keeps its prefix / suffix products
    assert n_slots <= (10 if kind == "perm" else 24)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,args", CASES + [("perm", (8, 10, 2, 3, 3, 7, 5)), ("logup", (9, 11, 2, 4))])
def test_h_term_programs_match_the_upstream_loops_on_device(ctx, kind, args):
    import torch

    c = build(kind, args)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()
    g = ctx.graph(c["calcs"], c["constants"], c["rotations"])
    vals = dev(c["prev"])
    fixed, advice = [dev(x) for x in c["fixed"]], [dev(x) for x in c["advice"]]
    torch.cuda.synchronize()
    g.evaluate(vals, c["ek"], c["rot_scale"], fixed=fixed, advice=advice, beta=c["beta"], gamma=c["gamma"], y=c["y"], extended_omega=c["we"])
    ctx.synchronize()
    assert np.array_equal(vals.cpu().numpy().view(np.uint64), c["want"])
    g.release()
