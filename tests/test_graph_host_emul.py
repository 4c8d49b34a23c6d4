"""The product's GraphEvaluator lowering (csrc/graph.hpp) and row interpreter (csrc/graph_exec.cuh) under host emulation,
checked against the oracle on random programs (CPU only; the -m gpu tests run the same programs through the CUDA kernel)."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from quotient_programs import (C_ADD, C_HORNER, C_MUL, C_STORE, R_MOD, S_ADVICE, S_CONST, S_INTER, S_PREV, omega_of, random_program)

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_emul", "ff_host.cpp")
SO = os.path.join(HERE, "host_emul", "libff_host.so")
CSRC = os.path.join(HERE, "..", "scroll-prover_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    deps = [SRC] + [os.path.join(CSRC, h) for h in ("ff.cuh", "ec.cuh", "graph.hpp", "graph_exec.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", SO, SRC])
    return C.CDLL(SO)


def table(cols):
    cols = [np.ascontiguousarray(c, dtype=np.uint64) for c in cols]
    return (C.c_void_p * max(1, len(cols)))(*[c.ctypes.data for c in cols]), cols


def host_eval(lib, calcs, constants, rotations, fixed, advice, instance, challenges, bgty, ext_omega, values, log_size, rot_scale):
    arr, parr, n_parts = O.pack_program(calcs)  # same struct layout as b200zk_calculation / b200zk_value_source
    consts = np.ascontiguousarray(np.asarray(constants, dtype=np.uint64).reshape(-1, 4))
    rots = np.ascontiguousarray(np.asarray(rotations, dtype=np.int32))
    tf, kf = table(fixed)
    ta, ka = table(advice)
    ti, ki = table(instance)
    ch = np.ascontiguousarray(np.asarray(challenges, dtype=np.uint64).reshape(-1, 4))
    sc = np.ascontiguousarray(np.stack(bgty).astype(np.uint64))
    out = np.ascontiguousarray(values, dtype=np.uint64).copy()
    info = (C.c_uint32 * 2)()
    err = C.create_string_buffer(256)
    vp = C.c_void_p
    eo = None if ext_omega is None else np.ascontiguousarray(ext_omega, dtype=np.uint64)
    rc = lib.graph_host_eval(arr, len(calcs), parr, n_parts, vp(consts.ctypes.data), len(consts), vp(rots.ctypes.data), len(rots),
                             tf, len(fixed), ta, len(advice), ti, len(instance), vp(ch.ctypes.data), len(ch), vp(sc.ctypes.data),
                             None if eo is None else vp(eo.ctypes.data), vp(out.ctypes.data), log_size, rot_scale, info, err)
    return rc, out, (info[0], info[1]), err.value.decode()


def inputs(seed, size, nf, na, ni, nc):
    rng = random.Random(seed)
    mk = lambda cnt: [O.fill_fr(size, rng.randrange(1 << 30)) for _ in range(cnt)]
    ch = O.fill_fr(max(nc, 1), 77)[:nc]
    bgty = [O.fill_fr(1, 1000 + i)[0] for i in range(4)]
    return mk(nf), mk(na), mk(ni), ch, bgty, O.fill_fr(size, 4242)


@pytest.mark.parametrize("seed,n_calcs,log_size,rot_scale,bias", [(1, 30, 4, 1, 0.5), (2, 200, 5, 4, 0.7), (3, 500, 3, 2, 0.3),
                                                                  (4, 64, 0, 1, 0.5), (5, 1000, 4, 4, 0.9), (6, 7, 6, 1, 0.0)])
def test_lowered_program_matches_oracle(lib, seed, n_calcs, log_size, rot_scale, bias):
    size = 1 << log_size
    calcs, constants, rotations = random_program(seed, n_calcs, 2, 4, 1, 3, 6, chain_bias=bias)
    fx, ad, ins, ch, bgty, prev = inputs(seed, size, 2, 4, 1, 3)
    w = O.fr_from_int(omega_of(log_size))
    consts = O.frs_from_ints(constants)
    want = O.graph_evaluate(calcs, consts, rotations, fx, ad, ins, ch, *bgty, w, prev, log_size, rot_scale)
    rc, got, (n_instr, n_slots), err = host_eval(lib, calcs, consts, rotations, fx, ad, ins, ch, bgty, w, prev, log_size, rot_scale)
    assert rc == 0, err
    assert np.array_equal(got, want)
    assert 2 <= n_slots <= 2 + n_calcs and n_instr >= 1


def test_slots_are_reused_and_dead_code_is_dropped(lib):
    # a long chain t_{i+1} = t_i * a + t_i only ever keeps two values live; 50 unused calculations emit nothing
    calcs = [(C_STORE, (S_ADVICE, 0, 0), None, None)]
    for i in range(300):
        calcs.append((C_MUL, (S_INTER, len(calcs) - 1, 0), (S_ADVICE, 0, 0), None))
        calcs.append((C_ADD, (S_INTER, len(calcs) - 1, 0), (S_INTER, len(calcs) - 2, 0), None))
    live = len(calcs)
    dead = [(C_MUL, (S_ADVICE, 0, 0), (S_ADVICE, 0, 0), None) for _ in range(50)]
    prog = calcs[:-1] + dead + [calcs[-1]]
    # re-index the final calculation's operands (they still name calcs[live-2] and calcs[live-3])
    ad = [O.fill_fr(8, 5)]
    z = O.fr_from_int(0)
    want = O.graph_evaluate(prog, np.zeros((0, 4), np.uint64), [0], [], ad, [], np.zeros((0, 4), np.uint64), z, z, z, z, None,
                            np.zeros((8, 4), np.uint64), 3, 1)
    rc, got, (n_instr, n_slots), err = host_eval(lib, prog, np.zeros((0, 4), np.uint64), [0], [], ad, [], np.zeros((0, 4), np.uint64),
                                                 [z, z, z, z], None, np.zeros((8, 4), np.uint64), 3, 1)
    assert rc == 0, err
    assert np.array_equal(got, want)
    assert n_instr == live and n_slots <= 2 + 3


def test_horner_destination_never_aliases_a_live_part(lib):
    # parts and factor die at the Horner itself: the destination must still be a different slot
    calcs = [(C_STORE, (S_ADVICE, 0, 0), None, None), (C_STORE, (S_ADVICE, 1, 0), None, None), (C_STORE, (S_ADVICE, 2, 0), None, None),
             (C_HORNER, (S_INTER, 0, 0), (S_INTER, 1, 0), [(S_INTER, 2, 0), (S_INTER, 0, 0), (S_INTER, 1, 0), (S_INTER, 2, 0)])]
    ad = [O.fill_fr(4, s) for s in (1, 2, 3)]
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    want = O.graph_evaluate(calcs, e, [0], [], ad, [], e, z, z, z, z, None, np.zeros((4, 4), np.uint64), 2, 1)
    rc, got, _, err = host_eval(lib, calcs, e, [0], [], ad, [], e, [z, z, z, z], None, np.zeros((4, 4), np.uint64), 2, 1)
    assert rc == 0, err
    assert np.array_equal(got, want)


def test_previous_value_chains_programs(lib):
    calcs = [(C_HORNER, (S_PREV, 0, 0), (S_CONST, 0, 0), [(S_ADVICE, 0, 0)])]
    ad = [O.fill_fr(16, 9)]
    consts = O.frs_from_ints([12345])
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    v_or = v_he = O.fill_fr(16, 10)
    for _ in range(3):
        v_or = O.graph_evaluate(calcs, consts, [0], [], ad, [], e, z, z, z, z, None, v_or, 4, 1)
        rc, v_he, _, err = host_eval(lib, calcs, consts, [0], [], ad, [], e, [z, z, z, z], None, v_he, 4, 1)
        assert rc == 0, err
    assert np.array_equal(v_or, v_he)


@pytest.mark.parametrize("calcs,msg", [
    ([(C_ADD, (S_INTER, 0, 0), (S_CONST, 0, 0), None)], "earlier calculation"),
    ([(C_ADD, (S_CONST, 5, 0), (S_CONST, 0, 0), None)], "constant index"),
    ([(C_ADD, (S_ADVICE, 0, 3), (S_CONST, 0, 0), None)], "rotation index"),
    ([(99, (S_CONST, 0, 0), None, None)], "unknown calculation"),
    ([(C_STORE, (42, 0, 0), None, None)], "unknown value source"),
])
def test_malformed_programs_are_rejected(lib, calcs, msg):
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    rc, _, _, err = host_eval(lib, calcs, O.frs_from_ints([1]), [0], [], [O.fill_fr(2, 1)], [], e, [z, z, z, z], None,
                              np.zeros((2, 4), np.uint64), 1, 1)
    assert rc == -1 and msg in err


def test_very_deep_dependency_chain(lib):
    # 60 000 calculations, each reading the previous one: the lowering walks it with an explicit stack
    calcs = [(C_STORE, (S_ADVICE, 0, 0), None, None)]
    for i in range(60000):
        calcs.append((C_MUL if i % 2 else C_ADD, (S_INTER, i, 0), (S_ADVICE, 0, 0), None))
    ad = [O.fill_fr(2, 5)]
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    want = O.graph_evaluate(calcs, e, [0], [], ad, [], e, z, z, z, z, None, np.zeros((2, 4), np.uint64), 1, 1)
    rc, got, (n_instr, n_slots), err = host_eval(lib, calcs, e, [0], [], ad, [], e, [z, z, z, z], None, np.zeros((2, 4), np.uint64), 1, 1)
    assert rc == 0, err
    assert np.array_equal(got, want) and n_instr == 60001 and n_slots == 3


def test_values_are_computed_right_before_their_first_reader(lib):
    # 300 gate values folded by one Horner: demand-driven emission keeps one of them live at a time
    calcs = [(C_MUL, (S_ADVICE, 0, 0), (S_ADVICE, 0, 1), None) for _ in range(300)]
    calcs.append((C_HORNER, (S_PREV, 0, 0), (S_CONST, 0, 0), [(S_INTER, i, 0) for i in range(300)]))
    ad = [O.fill_fr(8, 5)]
    consts = O.frs_from_ints([77])
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    prev = O.fill_fr(8, 6)
    want = O.graph_evaluate(calcs, consts, [0, 1], [], ad, [], e, z, z, z, z, None, prev, 3, 1)
    rc, got, (n_instr, n_slots), err = host_eval(lib, calcs, consts, [0, 1], [], ad, [], e, [z, z, z, z], None, prev, 3, 1)
    assert rc == 0, err
    assert np.array_equal(got, want)
    assert n_slots <= 2 + 2 and n_instr == 300 + 1 + 300  # 300 MUL, MOV, 300 MAD


def test_too_many_live_intermediates_is_reported(lib):
    # 300 values, each read by two Horner runs in opposite orders: all of them are live when the first run ends
    calcs = [(C_STORE, (S_ADVICE, 0, 0), None, None) for _ in range(300)]
    calcs.append((C_HORNER, (S_CONST, 0, 0), (S_CONST, 0, 0), [(S_INTER, i, 0) for i in range(300)]))
    calcs.append((C_HORNER, (S_CONST, 0, 0), (S_CONST, 0, 0), [(S_INTER, 299 - i, 0) for i in range(300)]))
    calcs.append((C_ADD, (S_INTER, 300, 0), (S_INTER, 301, 0), None))
    z = O.fr_from_int(0)
    e = np.zeros((0, 4), np.uint64)
    rc, _, _, err = host_eval(lib, calcs, O.frs_from_ints([1]), [0], [], [O.fill_fr(2, 1)], [], e, [z, z, z, z], None,
                              np.zeros((2, 4), np.uint64), 1, 1)
    assert rc == -1 and "too many intermediates" in err


def test_many_random_small_programs(lib):
    """300 random programs of every shape (operand sharing, Horner with repeated / empty parts, dead code, results that are
    plain sources): the lowered program must equal the oracle's straight evaluation, row for row."""
    rng = random.Random(20260922)
    for trial in range(300):
        n_calcs = rng.randrange(1, 70)
        bias = rng.choice([0.0, 0.2, 0.5, 0.8, 0.95])
        log_size = rng.choice([0, 1, 2, 3])
        rot_scale = rng.choice([1, 2, 4])
        calcs, constants, rotations = random_program(1000 + trial, n_calcs, 1, 2, 1, 1, rng.randrange(1, 5), n_constants=3,
                                                     use_prev=rng.random() < 0.7, use_x=rng.random() < 0.5, chain_bias=bias)
        size = 1 << log_size
        fx, ad, ins, ch, bgty, prev = inputs(trial, size, 1, 2, 1, 1)
        w = O.fr_from_int(omega_of(log_size))
        consts = O.frs_from_ints(constants)
        want = O.graph_evaluate(calcs, consts, rotations, fx, ad, ins, ch, *bgty, w, prev, log_size, rot_scale)
        rc, got, (n_instr, n_slots), err = host_eval(lib, calcs, consts, rotations, fx, ad, ins, ch, bgty, w, prev, log_size, rot_scale)
        assert rc == 0, (trial, err)
        assert np.array_equal(got, want), trial
        assert n_slots <= 2 + n_calcs
