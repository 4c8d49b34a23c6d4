"""evaluate_h's hard-coded sections written as GraphEvaluator programs (neutral form of quotient_programs.py).

Upstream (halo2_proofs @ e5ddf67, plonk/evaluation.rs) evaluates custom gates through GraphEvaluator but spells the
permutation and lookup identities out as Rust loops.  On the device both become programs (B200ZK_SRC_EXTENDED_X supplies the
coset point), so one kernel serves the whole of evaluate_h.  The C++ twin of these generators is
halo2_b200::plonk::{permutation_constraints, lookup_constraints} in scroll-prover_b200/halo2_b200.hpp.
"""
from quotient_programs import (C_ADD, C_HORNER, C_MUL, C_SQUARE, C_SUB, DELTA, R_MOD, S_ADVICE, S_BETA, S_CONST, S_FIXED, S_GAMMA,
                               S_INTER, S_PREV, S_X, S_Y)


class _Builder:
    def __init__(self):
        self.calcs = []

    def add(self, op, a, b=None, parts=None):
        self.calcs.append((op, a, b, parts))
        return (S_INTER, len(self.calcs) - 1, 0)


def permutation_terms_program(n_sets: int, chunk_len: int, n_cols: int, last_rotation: int):
    """The `// Permutations` section of evaluate_h.

    Column tables expected by the program:  advice = [z_0 .. z_{S-1}, v_0 .. v_{C-1}] (extended cosets),
    fixed = [sigma_0 .. sigma_{C-1}, l0, l_last, l_active_row];  rotations = [0, 1, last_rotation];
    constants = [0, 1, DELTA^0 .. DELTA^{C-1}].  Returns (calcs, constants, rotations)."""
    b = _Builder()
    one = (S_CONST, 1, 0)
    z = lambda s, r=0: (S_ADVICE, s, r)
    v = lambda j: (S_ADVICE, n_sets + j, 0)
    sig = lambda j: (S_FIXED, j, 0)
    l0, l_last, l_act = (S_FIXED, n_cols, 0), (S_FIXED, n_cols + 1, 0), (S_FIXED, n_cols + 2, 0)
    beta, gamma = (S_BETA, 0, 0), (S_GAMMA, 0, 0)
    terms = []
    t = b.add(C_SUB, one, z(0))
    terms.append(b.add(C_MUL, t, l0))                       # l_0 * (1 - z_0)
    t = b.add(C_SQUARE, z(n_sets - 1))
    t = b.add(C_SUB, t, z(n_sets - 1))
    terms.append(b.add(C_MUL, t, l_last))                   # l_last * (z_l^2 - z_l)
    for s in range(1, n_sets):
        t = b.add(C_SUB, z(s), z(s - 1, 2))
        terms.append(b.add(C_MUL, t, l0))                   # l_0 * (z_i - z_{i-1}(w^last X))
    bx = b.add(C_MUL, beta, (S_X, 0, 0))                    # delta_start * beta_term = beta * zeta * w_ext^idx
    for s in range(n_sets):
        cols = range(s * chunk_len, min((s + 1) * chunk_len, n_cols))
        left = z(s, 1)
        for j in cols:
            u = b.add(C_MUL, beta, sig(j))
            u = b.add(C_ADD, u, v(j))
            u = b.add(C_ADD, u, gamma)
            left = b.add(C_MUL, left, u)
        right = z(s)
        for j in cols:
            d = bx if j == 0 else b.add(C_MUL, bx, (S_CONST, 2 + j, 0))  # current_delta = beta * X * DELTA^j
            u = b.add(C_ADD, v(j), d)
            u = b.add(C_ADD, u, gamma)
            right = b.add(C_MUL, right, u)
        t = b.add(C_SUB, left, right)
        terms.append(b.add(C_MUL, t, l_act))
    b.add(C_HORNER, (S_PREV, 0, 0), (S_Y, 0, 0), terms)     # value = value*y + term, in upstream's order
    constants = [0, 1] + [pow(DELTA, j, R_MOD) for j in range(n_cols)]
    return b.calcs, constants, [0, 1, last_rotation]


def logup_terms_program(n_inputs: int):
    """The lookup section of evaluate_h for one log-derivative lookup.

    advice = [f_0 .. f_{m-1} (compressed inputs), t (compressed table), m, phi] on the extended coset,
    fixed = [l0, l_last, l_active_row]; rotations = [0, 1].  Upstream computes  rhs = prod * (tau * sum_i 1/phi_i - m)  with a
    batch inversion; the program uses the equal polynomial form  tau * sum_i prod_{j != i} phi_j - m * prod  (no inversion; the
    two differ only if some f_i + beta vanishes at a coset point, where upstream's value is not the polynomial's either)."""
    b = _Builder()
    beta = (S_BETA, 0, 0)
    f = lambda i: (S_ADVICE, i, 0)
    table, m = (S_ADVICE, n_inputs, 0), (S_ADVICE, n_inputs + 1, 0)
    phi = lambda r=0: (S_ADVICE, n_inputs + 2, r)
    l0, l_last, l_act = (S_FIXED, 0, 0), (S_FIXED, 1, 0), (S_FIXED, 2, 0)
    ph = [b.add(C_ADD, f(i), beta) for i in range(n_inputs)]
    pre = [None] * n_inputs   # prod_{j < i}
    suf = [None] * n_inputs   # prod_{j > i}
    acc = None
    for i in range(n_inputs):
        pre[i] = acc
        acc = ph[i] if acc is None else b.add(C_MUL, acc, ph[i])
    prod = acc if acc is not None else (S_CONST, 1, 0)
    acc = None
    for i in reversed(range(n_inputs)):
        suf[i] = acc
        acc = ph[i] if acc is None else b.add(C_MUL, acc, ph[i])
    ssum = None               # sum_i prod_{j != i} phi_j
    for i in range(n_inputs):
        if pre[i] is None and suf[i] is None:
            term = (S_CONST, 1, 0)
        elif pre[i] is None:
            term = suf[i]
        elif suf[i] is None:
            term = pre[i]
        else:
            term = b.add(C_MUL, pre[i], suf[i])
        ssum = term if ssum is None else b.add(C_ADD, ssum, term)
    if ssum is None:
        ssum = (S_CONST, 0, 0)
    tau = b.add(C_ADD, table, beta)
    d = b.add(C_SUB, phi(1), phi())
    lhs = b.add(C_MUL, b.add(C_MUL, tau, prod), d)
    rhs = b.add(C_SUB, b.add(C_MUL, tau, ssum), b.add(C_MUL, m, prod))
    q = b.add(C_MUL, b.add(C_SUB, lhs, rhs), l_act)
    t0 = b.add(C_MUL, l0, phi())
    t1 = b.add(C_MUL, l_last, phi())
    b.add(C_HORNER, (S_PREV, 0, 0), (S_Y, 0, 0), [t0, t1, q])
    return b.calcs, [0, 1], [0, 1]
