"""b200zk_graph_check: the product library's own program validation + lowering, callable without a GPU (a key-generation
host can vet a circuit's GraphEvaluator programs and read their on-chip footprint).  CPU only: loads libb200zk.so, no context."""
import pytest

from h_terms_programs import logup_terms_program, permutation_terms_program
from quotient_programs import C_ADD, C_HORNER, C_MUL, C_STORE, S_ADVICE, S_CONST, S_INTER, S_PREV, S_Y, random_program


def test_counts_for_known_shapes(zk):
    # y-fold over 300 gate values: 300 MUL + MOV + 300 MAD, a handful of slots
    calcs = [(C_MUL, (S_ADVICE, 0, 0), (S_ADVICE, 0, 1), None) for _ in range(300)]
    calcs.append((C_HORNER, (S_PREV, 0, 0), (S_CONST, 0, 0), [(S_INTER, i, 0) for i in range(300)]))
    info = zk.graph_check(calcs, 1, 2)
    assert info == {"n_instructions": 601, "n_slots": info["n_slots"]} and info["n_slots"] <= 4
    assert zk.graph_check([], 0, 1) == {"n_instructions": 0, "n_slots": 2}


@pytest.mark.parametrize("seed", range(5))
def test_random_programs_lower(zk, seed):
    calcs, constants, rotations = random_program(seed, 400, 2, 4, 1, 3, 6, chain_bias=0.6)
    info = zk.graph_check(calcs, len(constants), len(rotations))
    assert 1 <= info["n_instructions"] and 2 <= info["n_slots"] <= 224


def test_h_term_programs_footprint(zk):
    for n_sets, chunk, n_cols in ((1, 3, 2), (3, 3, 8), (7, 3, 20)):
        calcs, constants, rotations = permutation_terms_program(n_sets, chunk, n_cols, -6)
        assert zk.graph_check(calcs, len(constants), len(rotations))["n_slots"] <= 10
    for n_inputs in (1, 3, 6):
        calcs, constants, rotations = logup_terms_program(n_inputs)
        assert zk.graph_check(calcs, len(constants), len(rotations))["n_slots"] <= 24


@pytest.mark.parametrize("calcs,code,msg", [
    ([(C_ADD, (S_INTER, 0, 0), (S_CONST, 0, 0), None)], -1, "earlier calculation"),
    ([(C_ADD, (S_CONST, 5, 0), (S_CONST, 0, 0), None)], -1, "constant index"),
    ([(C_ADD, (S_ADVICE, 0, 9), (S_CONST, 0, 0), None)], -1, "rotation index"),
    ([(77, (S_CONST, 0, 0), None, None)], -1, "unknown calculation"),
])
def test_rejections_carry_the_reason(zk, calcs, code, msg):
    with pytest.raises(zk.B200zkError) as ei:
        zk.graph_check(calcs, 1, 1)
    assert ei.value.code == code and msg in str(ei.value)


def test_too_many_live_values_is_unsupported(zk):
    calcs = [(C_STORE, (S_ADVICE, 0, 0), None, None) for _ in range(300)]
    calcs.append((C_HORNER, (S_CONST, 0, 0), (S_Y, 0, 0), [(S_INTER, i, 0) for i in range(300)]))
    calcs.append((C_HORNER, (S_CONST, 0, 0), (S_Y, 0, 0), [(S_INTER, 299 - i, 0) for i in range(300)]))
    calcs.append((C_ADD, (S_INTER, 300, 0), (S_INTER, 301, 0), None))
    with pytest.raises(zk.B200zkError) as ei:
        zk.graph_check(calcs, 1, 1)
    assert ei.value.code == zk.E_UNSUPPORTED
