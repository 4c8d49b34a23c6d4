"""The permutation and log-derivative lookup arguments end to end on a small satisfied instance (CPU, oracle):

  witness with copy constraints / lookups  ->  z(X) per column set (permutation::Argument::commit, chained through
  z[u], blinding rows)  /  phi(X) (mv-lookup running sum)  ->  everything to the extended coset  ->  evaluate_h's
  permutation / lookup sections as GENERATED programs (tests/h_terms_programs.py)  ->  divide by X^n - 1  ->
  extended_to_coeff  ->  h(X)

and then what a verifier does: pick a random x, evaluate every polynomial at x (and at the rotated points) from its
coefficients, recompute the constraint expression with plain integers and check  expression(x) = h(x) * (x^n - 1).
This ties the restated prover steps and the generated programs to the argument's actual meaning (a quotient exists iff the
copy constraints / lookups hold) instead of to each other: a wrong rotation, chunk boundary, delta power, blinding offset or
sign anywhere makes the point check fail.  A broken witness must make it fail too.
"""
import random

import numpy as np
import pytest

from h_terms_programs import logup_terms_program, permutation_terms_program
from oracle import oracle as O
from quotient_programs import DELTA, R_MOD, ZETA, omega_of

K, EXT = 5, 2
N, EK = 1 << K, K + EXT
BLINDING = 3
U = N - (BLINDING + 1)          # rows [0, U) are usable, row U is "last", rows > U are blinding rows
LAST_ROTATION = -(BLINDING + 1)
W, WE = omega_of(K), omega_of(EK)
Z0 = O.fr_from_int(0)
E = np.zeros((0, 4), np.uint64)


def fr(v):
    return O.fr_from_int(v % R_MOD)


def lagrange_selectors():
    l0 = [1 if i == 0 else 0 for i in range(N)]
    l_last = [1 if i == U else 0 for i in range(N)]
    l_active = [1 if i < U else 0 for i in range(N)]  # 1 - (l_last + l_blind)
    return l0, l_last, l_active


class Domain:
    def __init__(self):
        self.d = O.EvaluationDomain(EXT + 3, K)  # j = 5: quotient degree 4, extended_k = k + 2
        assert self.d.extended_k == EK
        t = [pow((pow(ZETA, N, R_MOD) * pow(WE, N * i, R_MOD) - 1) % R_MOD, -1, R_MOD) for i in range(1 << EXT)]
        self.t_inv = [t[i % (1 << EXT)] for i in range(1 << EK)]

    def coeff(self, values):
        return self.d.lagrange_to_coeff(O.frs_from_ints(values))

    def extended(self, coeff):
        return self.d.coeff_to_extended(coeff)

    def quotient(self, numerator_ext):
        ints = O.frs_to_ints(numerator_ext)
        q = O.frs_from_ints([a * b % R_MOD for a, b in zip(ints, self.t_inv)])
        return self.d.extended_to_coeff(q)


def ev(coeff, x):
    return O.fr_to_int(O.eval_polynomial(coeff, fr(x)))


def permutation_instance(seed, n_cols, chunk_len, break_it=False):
    rng = random.Random(seed)
    cells = [(j, i) for j in range(n_cols) for i in range(U)]  # copy constraints live on the usable rows
    perm = list(range(len(cells)))
    rng.shuffle(perm)
    label = lambda j, i: pow(DELTA, j, R_MOD) * pow(W, i, R_MOD) % R_MOD
    vals = [[rng.randrange(R_MOD) for _ in range(N)] for _ in range(n_cols)]   # blinding rows: random
    sig = [[label(j, i) for i in range(N)] for j in range(n_cols)]            # identity outside the cycles
    pos = 0
    while pos < len(perm):
        grp = perm[pos:pos + rng.randrange(1, 5)]
        pos += len(grp)
        v = rng.randrange(R_MOD)
        for t, ci in enumerate(grp):
            j, i = cells[ci]
            vals[j][i] = v
            nj, ni = cells[grp[(t + 1) % len(grp)]]
            sig[j][i] = label(nj, ni)
    if break_it:
        j, i = cells[perm[0]] if len(perm) else (0, 0)
        # find a cell that is tied to a different cell and change its value only
        for ci in perm:
            j, i = cells[ci]
            if sig[j][i] != label(j, i):
                vals[j][i] = (vals[j][i] + 1) % R_MOD
                break
    return vals, sig


@pytest.mark.parametrize("n_cols,chunk_len,break_it", [(5, 3, False), (3, 3, False), (4, 2, False), (5, 3, True)])
def test_permutation_argument_quotient_identity(n_cols, chunk_len, break_it):
    dom = Domain()
    rng = random.Random(1000 + n_cols)
    beta, gamma, y = (rng.randrange(R_MOD) for _ in range(3))
    vals, sig = permutation_instance(7 + n_cols, n_cols, chunk_len, break_it)
    n_sets = (n_cols + chunk_len - 1) // chunk_len
    # ---- permutation::Argument::commit: one z per column chunk, chained through z[U], blinding rows random
    zs, last_z, dws = [], 1, 1
    for s in range(n_sets):
        cols = range(s * chunk_len, min((s + 1) * chunk_len, n_cols))
        z = O.frs_to_ints(O.permutation_product([O.frs_from_ints(vals[j]) for j in cols], [O.frs_from_ints(sig[j]) for j in cols],
                                                fr(beta), fr(gamma), fr(dws), fr(DELTA), fr(W), K, fr(last_z)))
        for i in range(N - BLINDING, N):
            z[i] = rng.randrange(R_MOD)
        last_z = z[U]
        dws = dws * pow(DELTA, len(cols), R_MOD) % R_MOD
        zs.append(z)
    if not break_it:
        assert zs[0][0] == 1 and zs[-1][U] == 1  # the grand product closes over the usable rows
    l0, l_last, l_act = lagrange_selectors()
    coeff = {"z": [dom.coeff(z) for z in zs], "v": [dom.coeff(v) for v in vals], "s": [dom.coeff(s) for s in sig],
             "l0": dom.coeff(l0), "l_last": dom.coeff(l_last), "l_act": dom.coeff(l_act)}
    ext = lambda c: dom.extended(c)
    calcs, constants, rotations = permutation_terms_program(n_sets, chunk_len, n_cols, LAST_ROTATION)
    advice = [ext(c) for c in coeff["z"]] + [ext(c) for c in coeff["v"]]
    fixed = [ext(c) for c in coeff["s"]] + [ext(coeff["l0"]), ext(coeff["l_last"]), ext(coeff["l_act"])]
    num = O.graph_evaluate(calcs, O.frs_from_ints(constants), rotations, fixed, advice, [], E, fr(beta), fr(gamma), Z0, fr(y), fr(WE),
                           np.zeros((1 << EK, 4), np.uint64), EK, 1 << EXT)
    h = dom.quotient(num)
    # ---- the verifier's side: recompute the expression at a random point with plain integers
    x = rng.randrange(R_MOD)
    xn, xl = x * W % R_MOD, x * pow(W, LAST_ROTATION % N, R_MOD) % R_MOD
    zx = [ev(c, x) for c in coeff["z"]]
    zxn = [ev(c, xn) for c in coeff["z"]]
    zxl = [ev(c, xl) for c in coeff["z"]]
    vx = [ev(c, x) for c in coeff["v"]]
    sx = [ev(c, x) for c in coeff["s"]]
    l0x, llx, lax = ev(coeff["l0"], x), ev(coeff["l_last"], x), ev(coeff["l_act"], x)
    terms = [(1 - zx[0]) * l0x, (zx[-1] * zx[-1] - zx[-1]) * llx]
    terms += [(zx[s] - zxl[s - 1]) * l0x for s in range(1, n_sets)]
    for s in range(n_sets):
        left, right = zxn[s], zx[s]
        for j in range(s * chunk_len, min((s + 1) * chunk_len, n_cols)):
            left = left * (vx[j] + beta * sx[j] + gamma) % R_MOD
            right = right * (vx[j] + pow(DELTA, j, R_MOD) * beta * x + gamma) % R_MOD
        terms.append((left - right) * lax)
    expr = 0
    for t in terms:
        expr = (expr * y + t) % R_MOD
    holds = expr == ev(h, x) * (pow(x, N, R_MOD) - 1) % R_MOD
    assert holds != break_it


# numerator degree (n_inputs + 3)(n - 1) must stay below 5n for the 4n-point extended domain: at most 2 inputs per lookup here
@pytest.mark.parametrize("n_inputs,break_it", [(1, False), (2, False), (2, True)])
def test_lookup_argument_quotient_identity(n_inputs, break_it):
    dom = Domain()
    rng = random.Random(2000 + n_inputs)
    beta, y = rng.randrange(R_MOD), rng.randrange(R_MOD)
    table = rng.sample(range(1, 100000), N)
    inputs = [[rng.choice(table[:U]) for _ in range(N)] for _ in range(n_inputs)]
    m = [sum(col[:U].count(t) for col in inputs) if i < U else 0 for i, t in enumerate(table)]  # multiplicities over usable rows
    if break_it:
        inputs[0][3] = 100001  # not in the table
    phi = O.frs_to_ints(O.logup_running_sum([O.frs_from_ints(c) for c in inputs], O.frs_from_ints(table), O.frs_from_ints(m), fr(beta), K,
                                            fr(0)))
    if not break_it:
        assert phi[0] == 0 and phi[U] == 0  # the log-derivative sum closes over the usable rows
    for i in range(N - BLINDING, N):
        phi[i] = rng.randrange(R_MOD)
    l0, l_last, l_act = lagrange_selectors()
    cf = {"f": [dom.coeff(c) for c in inputs], "t": dom.coeff(table), "m": dom.coeff(m), "phi": dom.coeff(phi),
          "l0": dom.coeff(l0), "l_last": dom.coeff(l_last), "l_act": dom.coeff(l_act)}
    calcs, constants, rotations = logup_terms_program(n_inputs)
    advice = [dom.extended(c) for c in cf["f"]] + [dom.extended(cf["t"]), dom.extended(cf["m"]), dom.extended(cf["phi"])]
    fixed = [dom.extended(cf["l0"]), dom.extended(cf["l_last"]), dom.extended(cf["l_act"])]
    num = O.graph_evaluate(calcs, O.frs_from_ints(constants), rotations, fixed, advice, [], E, fr(beta), Z0, Z0, fr(y), None,
                           np.zeros((1 << EK, 4), np.uint64), EK, 1 << EXT)
    h = dom.quotient(num)
    x = rng.randrange(R_MOD)
    fx = [ev(c, x) for c in cf["f"]]
    tx, mx, px, pxn = ev(cf["t"], x), ev(cf["m"], x), ev(cf["phi"], x), ev(cf["phi"], x * W % R_MOD)
    l0x, llx, lax = ev(cf["l0"], x), ev(cf["l_last"], x), ev(cf["l_act"], x)
    prod = 1
    for f in fx:
        prod = prod * (f + beta) % R_MOD
    inv_sum = sum(pow(f + beta, -1, R_MOD) for f in fx) % R_MOD
    tau = (tx + beta) % R_MOD
    lhs = tau * prod * (pxn - px) % R_MOD
    rhs = prod * (tau * inv_sum - mx) % R_MOD
    expr = 0
    for t in (l0x * px, llx * px, (lhs - rhs) * lax):
        expr = (expr * y + t) % R_MOD
    holds = expr == ev(h, x) * (pow(x, N, R_MOD) - 1) % R_MOD
    assert holds != break_it
