"""Pins oracle/halo2_quotient.c (graph evaluator, permutation product, log-derivative running sum) against an
independent big-integer model and against the algebraic identities the arguments rest on (CPU only)."""
import random

import numpy as np
import pytest

from oracle import oracle as O
from quotient_programs import (DELTA, R_MOD, model_evaluate, model_logup, model_permutation_product, omega_of, random_program)


def cols_of(rng, count, size):
    ints = [[rng.randrange(R_MOD) for _ in range(size)] for _ in range(count)]
    return ints, [O.frs_from_ints(c) for c in ints]


@pytest.mark.parametrize("seed,log_size,rot_scale", [(1, 4, 1), (2, 5, 4), (3, 3, 2), (4, 0, 1), (5, 6, 4)])
def test_graph_evaluate_matches_bigint_model(seed, log_size, rot_scale):
    rng = random.Random(seed)
    size = 1 << log_size
    calcs, constants, rotations = random_program(seed, 40, 2, 3, 1, 2, 5)
    fx_i, fx = cols_of(rng, 2, size)
    ad_i, ad = cols_of(rng, 3, size)
    in_i, ins = cols_of(rng, 1, size)
    ch = [rng.randrange(R_MOD) for _ in range(2)]
    beta, gamma, theta, y = (rng.randrange(R_MOD) for _ in range(4))
    prev = [rng.randrange(R_MOD) for _ in range(size)]
    w = omega_of(log_size)
    want = model_evaluate(calcs, constants, rotations, fx_i, ad_i, in_i, ch, beta, gamma, theta, y, w, prev, log_size, rot_scale)
    got = O.graph_evaluate(calcs, O.frs_from_ints(constants), rotations, fx, ad, ins, O.frs_from_ints(ch), O.fr_from_int(beta),
                           O.fr_from_int(gamma), O.fr_from_int(theta), O.fr_from_int(y), O.fr_from_int(w), O.frs_from_ints(prev),
                           log_size, rot_scale)
    assert O.frs_to_ints(got) == want


def test_graph_evaluate_empty_program_is_zero():
    vals = O.frs_from_ints([5, 6, 7, 8])
    z = O.fr_from_int(0)
    got = O.graph_evaluate([], np.zeros((0, 4), np.uint64), [0], [], [], [], np.zeros((0, 4), np.uint64), z, z, z, z, None, vals, 2, 1)
    assert O.frs_to_ints(got) == [0, 0, 0, 0]


def test_gate_fold_with_y_is_horner_in_y():
    """evaluate_h folds gates as value = value * y + gate: two chained programs equal the two-term polynomial in y."""
    rng = random.Random(9)
    size = 8
    a_i, a = cols_of(rng, 2, size)
    y = rng.randrange(R_MOD)
    from quotient_programs import C_HORNER, C_MUL, C_SUB, S_ADVICE, S_INTER, S_PREV, S_Y
    g1 = [(C_MUL, (S_ADVICE, 0, 0), (S_ADVICE, 1, 0), None), (C_HORNER, (S_PREV, 0, 0), (S_Y, 0, 0), [(S_INTER, 0, 0)])]
    g2 = [(C_SUB, (S_ADVICE, 0, 0), (S_ADVICE, 1, 1), None), (C_HORNER, (S_PREV, 0, 0), (S_Y, 0, 0), [(S_INTER, 0, 0)])]
    z = O.fr_from_int(0)
    v = np.zeros((size, 4), np.uint64)
    for g in (g1, g2):
        v = O.graph_evaluate(g, np.zeros((0, 4), np.uint64), [0, 1], [], a, [], np.zeros((0, 4), np.uint64), z, z, z, O.fr_from_int(y),
                             None, v, 3, 1)
    want = [((a_i[0][i] * a_i[1][i]) * y + (a_i[0][i] - a_i[1][(i + 1) % size])) % R_MOD for i in range(size)]
    assert O.frs_to_ints(v) == want


@pytest.mark.parametrize("k,n_cols", [(3, 1), (5, 3), (6, 4), (0, 2)])
def test_permutation_product_matches_model(k, n_cols):
    rng = random.Random(100 + k)
    n = 1 << k
    v_i, v = cols_of(rng, n_cols, n)
    s_i, s = cols_of(rng, n_cols, n)
    beta, gamma, z0 = (rng.randrange(R_MOD) for _ in range(3))
    dws = pow(DELTA, 3, R_MOD)
    w = omega_of(k)
    want = model_permutation_product(v_i, s_i, beta, gamma, dws, DELTA, w, k, z0)
    got = O.permutation_product(v, s, O.fr_from_int(beta), O.fr_from_int(gamma), O.fr_from_int(dws), O.fr_from_int(DELTA),
                                O.fr_from_int(w), k, O.fr_from_int(z0))
    assert O.frs_to_ints(got) == want


def test_permutation_product_telescopes_for_a_valid_permutation():
    """sigma encodes a real permutation of the (column, row) cells that preserves values: z(omega^n) comes back to 1."""
    rng = random.Random(7)
    k, n_cols = 5, 3
    n = 1 << k
    w = omega_of(k)
    cells = [(j, i) for j in range(n_cols) for i in range(n)]
    # equality classes: random groups of cells share a value; sigma cycles each group
    perm = list(range(len(cells)))
    rng.shuffle(perm)
    groups, pos = [], 0
    while pos < len(perm):
        g = rng.randrange(1, 5)
        groups.append(perm[pos:pos + g])
        pos += g
    vals = [[0] * n for _ in range(n_cols)]
    label = lambda c: pow(DELTA, c[0], R_MOD) * pow(w, c[1], R_MOD) % R_MOD
    sig = [[0] * n for _ in range(n_cols)]
    for g in groups:
        val = rng.randrange(R_MOD)
        for t, ci in enumerate(g):
            j, i = cells[ci]
            vals[j][i] = val
            sig[j][i] = label(cells[g[(t + 1) % len(g)]])
    beta, gamma = rng.randrange(R_MOD), rng.randrange(R_MOD)
    z = O.permutation_product([O.frs_from_ints(c) for c in vals], [O.frs_from_ints(c) for c in sig], O.fr_from_int(beta),
                              O.fr_from_int(gamma), O.fr_from_int(1), O.fr_from_int(DELTA), O.fr_from_int(w), k, O.fr_from_int(1))
    zi = O.frs_to_ints(z)
    # last step closes the product: z[n-1] * mv[n-1] == 1
    num = den = 1
    for j in range(n_cols):
        num = num * (label((j, n - 1)) * beta + gamma + vals[j][n - 1]) % R_MOD
        den = den * (beta * sig[j][n - 1] + gamma + vals[j][n - 1]) % R_MOD
    assert zi[0] == 1 and zi[n - 1] * num % R_MOD * pow(den, -1, R_MOD) % R_MOD == 1


@pytest.mark.parametrize("k,n_inputs", [(4, 1), (5, 3), (0, 1)])
def test_logup_matches_model(k, n_inputs):
    rng = random.Random(200 + k)
    n = 1 << k
    f_i, f = cols_of(rng, n_inputs, n)
    t_i, t = cols_of(rng, 1, n)
    m_i = [rng.randrange(0, 5) for _ in range(n)]
    beta, p0 = rng.randrange(R_MOD), rng.randrange(R_MOD)
    if n > 2:
        f_i[0][1] = (-beta) % R_MOD  # a zero denominator stays zero under BatchInvert
        f[0] = O.frs_from_ints(f_i[0])
    want = model_logup(f_i, t_i[0], m_i, beta, k, p0)
    got = O.logup_running_sum(f, t[0], O.frs_from_ints(m_i), O.fr_from_int(beta), k, O.fr_from_int(p0))
    assert O.frs_to_ints(got) == want


def test_logup_telescopes_for_a_valid_lookup():
    """every input value occurs in the table and m counts the occurrences: the running sum closes at zero."""
    rng = random.Random(11)
    k, n_inputs = 5, 2
    n = 1 << k
    table = rng.sample(range(1, 1000), n)
    inputs = [[rng.choice(table) for _ in range(n)] for _ in range(n_inputs)]
    m = [sum(col.count(t) for col in inputs) for t in table]
    beta = rng.randrange(R_MOD)
    phi = O.frs_to_ints(O.logup_running_sum([O.frs_from_ints(c) for c in inputs], O.frs_from_ints(table), O.frs_from_ints(m),
                                            O.fr_from_int(beta), k, O.fr_from_int(0)))
    inv = lambda x: pow(x, -1, R_MOD)
    last = (sum(inv(c[n - 1] + beta) for c in inputs) - m[n - 1] * inv(table[n - 1] + beta)) % R_MOD
    assert phi[0] == 0 and (phi[n - 1] + last) % R_MOD == 0


def test_prefix_scan_both_ops():
    rng = random.Random(3)
    a = [rng.randrange(R_MOD) for _ in range(37)]
    init = rng.randrange(R_MOD)
    p = O.frs_to_ints(O.prefix_scan(0, O.frs_from_ints(a), O.fr_from_int(init)))
    s = O.frs_to_ints(O.prefix_scan(1, O.frs_from_ints(a), O.fr_from_int(init)))
    acc_p, acc_s = init, init
    for i in range(37):
        assert p[i] == acc_p and s[i] == acc_s
        acc_p = acc_p * a[i] % R_MOD
        acc_s = (acc_s + a[i]) % R_MOD
