"""Parity of the quotient-construction kernels (csrc/quotient.cu) with the oracle, through the C ABI, on device-resident
columns: GraphEvaluator programs over the extended domain, the permutation z(X) product and the log-derivative phi(X) sum.

Reference behaviour under test: halo2_proofs @ e5ddf67 plonk/evaluation.rs (GraphEvaluator::evaluate),
plonk/permutation/prover.rs (Argument::commit) and plonk/mv_lookup/prover.rs -- see oracle/halo2_quotient.c.
"""
import random

import numpy as np
import pytest

from oracle import oracle as O
from quotient_programs import (C_HORNER, C_MUL, C_STORE, C_SUB, DELTA, R_MOD, S_ADVICE, S_BETA, S_CONST, S_FIXED, S_GAMMA, S_INTER, S_PREV,
                               S_X, S_Y, ZETA, omega_of, random_program)

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def one_stream(ctx):
    """The ABI is asynchronous on the context stream for device-resident outputs.  These tests hand torch tensors in and
    read torch tensors back, so the library is put on torch's current (non-default) stream for their duration: clones,
    kernels and .cpu() copies are then ordered on one stream, as a device-resident prover session would run them."""
    import torch

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ctx.set_stream(s.cuda_stream)
        yield
        ctx.synchronize()
    ctx.set_stream(None)


def dev(a):
    import torch

    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def inputs(seed, size, nf, na, ni, nc):
    rng = random.Random(seed)
    mk = lambda cnt: [O.fill_fr(size, rng.randrange(1 << 30)) for _ in range(cnt)]
    ch = O.fill_fr(max(nc, 1), 77)[:nc]
    bgty = [O.fill_fr(1, 1000 + i)[0] for i in range(4)]
    return mk(nf), mk(na), mk(ni), ch, bgty, O.fill_fr(size, 4242)


@pytest.mark.parametrize("seed,n_calcs,log_size,rot_scale,bias", [(1, 30, 4, 1, 0.5), (2, 200, 10, 4, 0.7), (3, 500, 7, 2, 0.3),
                                                                  (4, 64, 0, 1, 0.5), (5, 1000, 12, 4, 0.9), (6, 7, 15, 1, 0.0),
                                                                  (7, 120, 5, 1, 0.2), (8, 2000, 9, 4, 0.1)])
def test_graph_evaluate_matches_oracle(ctx, zk, seed, n_calcs, log_size, rot_scale, bias):
    size = 1 << log_size
    calcs, constants, rotations = random_program(seed, n_calcs, 2, 4, 1, 3, 6, chain_bias=bias)
    fx, ad, ins, ch, bgty, prev = inputs(seed, size, 2, 4, 1, 3)
    w = O.fr_from_int(omega_of(log_size))
    consts = O.frs_from_ints(constants)
    want = O.graph_evaluate(calcs, consts, rotations, fx, ad, ins, ch, *bgty, w, prev, log_size, rot_scale)
    g = ctx.graph(calcs, consts, rotations)
    info = g.info()
    assert 2 <= info["n_slots"] <= 224
    vals = dev(prev)
    dfx, dad, dins = [dev(c) for c in fx], [dev(c) for c in ad], [dev(c) for c in ins]
    g.evaluate(vals, log_size, rot_scale, fixed=dfx, advice=dad, instance=dins, challenges=ch, beta=bgty[0], gamma=bgty[1],
               theta=bgty[2], y=bgty[3], extended_omega=w)
    assert np.array_equal(host(vals), want)
    # a second evaluation of the same handle chains on the previous values (PreviousValue), as evaluate_h does per gate
    want2 = O.graph_evaluate(calcs, consts, rotations, fx, ad, ins, ch, *bgty, w, want, log_size, rot_scale)
    g.evaluate(vals, log_size, rot_scale, fixed=dfx, advice=dad, instance=dins, challenges=ch, beta=bgty[0], gamma=bgty[1],
               theta=bgty[2], y=bgty[3], extended_omega=w)
    assert np.array_equal(host(vals), want2)
    g.release()


def test_graph_many_live_slots_uses_narrow_blocks(ctx):
    # 150 values read by two Horner runs in opposite orders stay live together: > 150 slots -> 32 rows per block
    n_live = 150
    calcs = [(C_MUL, (S_ADVICE, i % 3, 0), (S_ADVICE, (i + 1) % 3, 1), None) for i in range(n_live)]
    calcs.append((C_HORNER, (S_CONST, 0, 0), (S_Y, 0, 0), [(S_INTER, i, 0) for i in range(n_live)]))
    calcs.append((C_HORNER, (S_CONST, 0, 0), (S_Y, 0, 0), [(S_INTER, n_live - 1 - i, 0) for i in range(n_live)]))
    calcs.append((C_MUL, (S_INTER, n_live, 0), (S_INTER, n_live + 1, 0), None))
    log_size = 8
    ad = [O.fill_fr(1 << log_size, s) for s in (1, 2, 3)]
    consts = O.frs_from_ints([7])
    z = O.fr_from_int(0)
    y = O.fr_from_int(123456789)
    e = np.zeros((0, 4), np.uint64)
    want = O.graph_evaluate(calcs, consts, [0, 1], [], ad, [], e, z, z, z, y, None, np.zeros((1 << log_size, 4), np.uint64), log_size, 1)
    g = ctx.graph(calcs, consts, [0, 1])
    assert n_live <= g.info()["n_slots"] <= n_live + 4
    vals = dev(np.zeros((1 << log_size, 4), np.uint64))
    g.evaluate(vals, log_size, 1, advice=[dev(c) for c in ad], y=y)
    assert np.array_equal(host(vals), want)


def test_graph_rejects_bad_programs_and_arguments(ctx, zk):
    with pytest.raises(zk.B200zkError) as ei:
        ctx.graph([(C_SUB, (S_INTER, 3, 0), (S_CONST, 0, 0), None)], O.frs_from_ints([1]), [0])
    assert ei.value.code == zk.E_INVALID and "earlier calculation" in str(ei.value)
    too_live = [(C_STORE, (S_ADVICE, 0, 0), None, None) for _ in range(300)]
    too_live.append((C_HORNER, (S_CONST, 0, 0), (S_CONST, 0, 0), [(S_INTER, i, 0) for i in range(300)]))
    too_live.append((C_HORNER, (S_CONST, 0, 0), (S_CONST, 0, 0), [(S_INTER, 299 - i, 0) for i in range(300)]))
    too_live.append((0, (S_INTER, 300, 0), (S_INTER, 301, 0), None))
    with pytest.raises(zk.B200zkError) as ei:
        ctx.graph(too_live, O.frs_from_ints([1]), [0])
    assert ei.value.code == zk.E_UNSUPPORTED
    g = ctx.graph([(C_MUL, (S_ADVICE, 1, 0), (S_FIXED, 0, 0), None)], O.frs_from_ints([1]), [0])
    vals = dev(np.zeros((4, 4), np.uint64))
    with pytest.raises(zk.B200zkError):  # the program reads advice[1] and fixed[0]; only one advice column is supplied
        g.evaluate(vals, 2, 1, advice=[dev(O.fill_fr(4, 1))], fixed=[dev(O.fill_fr(4, 2))])
    with pytest.raises(zk.B200zkError):  # host memory is not accepted for columns
        g.evaluate(vals, 2, 1, advice=[O.fill_fr(4, 1), O.fill_fr(4, 3)], fixed=[dev(O.fill_fr(4, 2))])
    gx = ctx.graph([(C_STORE, (S_X, 0, 0), None, None)], O.frs_from_ints([1]), [0])
    with pytest.raises(zk.B200zkError):  # ExtendedX needs extended_omega
        gx.evaluate(vals, 2, 1)


def test_permutation_identity_as_a_program_vanishes_on_a_valid_witness(ctx):
    """End to end over the pieces: z(X) from b200zk_permutation_product, coset-extended with the NTT path, satisfies
    z(wX) * prod(v + beta*sigma + gamma) - z(X) * prod(v + delta^j*beta*X + gamma) = 0 on the extended coset, the identity
    evaluate_h encodes, here written as a GraphEvaluator program with ExtendedX: its values on the coset, divided by
    X^n - 1 and brought back with extended_to_coeff, must be a polynomial of degree < n_cols*(n-1) (exact division)."""
    rng = random.Random(21)
    k, n_cols, ext = 6, 2, 2
    n, ek = 1 << k, k + 2
    w, we = omega_of(k), omega_of(ek)
    cells = [(j, i) for j in range(n_cols) for i in range(n)]
    perm = list(range(len(cells)))
    rng.shuffle(perm)
    label = lambda c: pow(DELTA, c[0], R_MOD) * pow(w, c[1], R_MOD) % R_MOD
    vals = [[0] * n for _ in range(n_cols)]
    sig = [[0] * n for _ in range(n_cols)]
    pos = 0
    while pos < len(perm):
        grp = perm[pos:pos + rng.randrange(1, 4)]
        pos += len(grp)
        if len(grp) > 1 and cells[grp[0]][0] == 0:
            tied_row = cells[grp[0]][1]  # a column-0 cell that is copy-constrained to another cell
        val = rng.randrange(R_MOD)
        for t, ci in enumerate(grp):
            j, i = cells[ci]
            vals[j][i] = val
            sig[j][i] = label(cells[grp[(t + 1) % len(grp)]])
    beta, gamma = rng.randrange(R_MOD), rng.randrange(R_MOD)
    fr = O.fr_from_int
    dv = [dev(O.frs_from_ints(c)) for c in vals]
    ds = [dev(O.frs_from_ints(c)) for c in sig]
    z = dev(np.zeros((n, 4), np.uint64))
    ctx.permutation_product(dv, ds, fr(beta), fr(gamma), fr(1), fr(DELTA), fr(w), k, fr(1), z)
    want_z = O.permutation_product([O.frs_from_ints(c) for c in vals], [O.frs_from_ints(c) for c in sig], fr(beta), fr(gamma), fr(1),
                                   fr(DELTA), fr(w), k, fr(1))
    assert np.array_equal(host(z), want_z)
    # to the extended coset: lagrange_to_coeff then coeff_to_extended for z, the columns and the sigmas
    import torch

    def extend(col):
        c = col.clone()
        ctx.best_fft(c, fr(pow(w, -1, R_MOD)), k, inverse_scale=True)
        out = torch.empty((1 << ek, 4), dtype=torch.int64, device="cuda")
        ctx.ntt_ext(c, k, out, ek, fr(we), coset_mode=1)
        return out
    cols = [extend(z)] + [extend(c) for c in dv]
    sigs = [extend(c) for c in ds]
    # advice = [z, v0, v1], fixed = [sigma0, sigma1]; rotation 1 = next row of the ORIGINAL domain (rot_scale = 4)
    A = lambda i, r=0: (S_ADVICE, i, r)
    F = lambda i: (S_FIXED, i, 0)
    I = lambda i: (S_INTER, i, 0)
    prog = [
        (C_MUL, (S_BETA, 0, 0), F(0), None),                      # 0: beta*s0
        (C_HORNER, I(0), (S_CONST, 1, 0), [(S_GAMMA, 0, 0)]),     # 1: beta*s0 + gamma       (factor 1)
        (0, I(1), A(1), None),                                    # 2: + v0
        (C_MUL, (S_BETA, 0, 0), F(1), None),                      # 3
        (C_HORNER, I(3), (S_CONST, 1, 0), [(S_GAMMA, 0, 0)]),     # 4
        (0, I(4), A(2), None),                                    # 5
        (C_MUL, I(2), I(5), None),                                # 6: left product
        (C_MUL, I(6), A(0, 1), None),                             # 7: z(wX) * left
        (C_MUL, (S_BETA, 0, 0), (S_X, 0, 0), None),               # 8: beta*X
        (C_HORNER, I(8), (S_CONST, 1, 0), [(S_GAMMA, 0, 0)]),     # 9: beta*X + gamma
        (0, I(9), A(1), None),                                    # 10
        (C_MUL, I(8), (S_CONST, 2, 0), None),                     # 11: delta*beta*X
        (C_HORNER, I(11), (S_CONST, 1, 0), [(S_GAMMA, 0, 0)]),    # 12
        (0, I(12), A(2), None),                                   # 13
        (C_MUL, I(10), I(13), None),                              # 14: right product
        (C_MUL, I(14), A(0), None),                               # 15: z(X) * right
        (C_SUB, I(7), I(15), None),                               # 16
    ]
    consts = O.frs_from_ints([0, 1, DELTA])
    g = ctx.graph(prog, consts, [0, 1])
    # (X^n - 1)^-1 on the extended coset: (zeta * we^i)^n takes 2^ext values
    tinv = [pow((pow(ZETA, n, R_MOD) * pow(we, n * (i % (1 << ext)), R_MOD) - 1) % R_MOD, -1, R_MOD) for i in range(1 << ext)]
    tinv_col = dev(O.frs_from_ints([tinv[i % (1 << ext)] for i in range(1 << ek)]))

    def quotient_tail(z_col, v0_col):
        out = dev(np.zeros((1 << ek, 4), np.uint64))
        adv = [z_col, v0_col, cols[2]]
        g.evaluate(out, ek, 1 << ext, fixed=sigs, advice=adv, beta=fr(beta), gamma=fr(gamma), extended_omega=fr(we))
        z0 = fr(0)
        want = O.graph_evaluate(prog, consts, [0, 1], [host(c) for c in sigs], [host(c) for c in adv], [], np.zeros((0, 4), np.uint64),
                                fr(beta), fr(gamma), z0, z0, fr(we), np.zeros((1 << ek, 4), np.uint64), ek, 1 << ext)
        assert np.array_equal(host(out), want)
        ctx.poly_mul(out, tinv_col, out=out)                               # / (X^n - 1) on the coset
        ctx.best_fft(out, fr(pow(we, -1, R_MOD)), ek, inverse_scale=True, coset_mode=2)  # extended_to_coeff
        return host(out)

    # numerator degree <= (1 + n_cols)(n - 1): a valid witness divides exactly, h has degree <= n_cols*(n-1) - 1
    h = quotient_tail(cols[0], cols[1])
    assert not h[n_cols * (n - 1):].any() and h[: n_cols * (n - 1)].any()
    # and it is not vacuous: break one cell of the witness and the "quotient" is no longer a low-degree polynomial
    bad = dv[0].clone()
    bad[tied_row] = dev(O.frs_from_ints([12345]))[0]
    h_bad = quotient_tail(cols[0], extend(bad))
    assert h_bad[n_cols * (n - 1):].any()


@pytest.mark.parametrize("k,n_cols", [(0, 1), (3, 1), (9, 3), (13, 5), (16, 2)])
def test_permutation_product_matches_oracle(ctx, k, n_cols):
    n = 1 << k
    v = [O.fill_fr(n, 300 + 7 * j + k) for j in range(n_cols)]
    s = [O.fill_fr(n, 900 + 5 * j + k) for j in range(n_cols)]
    beta, gamma, z0 = O.fill_fr(3, 55 + k)
    dws = O.fr_from_int(pow(DELTA, 4, R_MOD))
    w = O.fr_from_int(omega_of(k))
    want = O.permutation_product(v, s, beta, gamma, dws, O.fr_from_int(DELTA), w, k, z0)
    out = dev(np.zeros((n, 4), np.uint64))
    ctx.permutation_product([dev(c) for c in v], [dev(c) for c in s], beta, gamma, dws, O.fr_from_int(DELTA), w, k, z0, out)
    assert np.array_equal(host(out), want)


@pytest.mark.parametrize("k,n_inputs", [(0, 1), (4, 1), (10, 3), (15, 2)])
def test_logup_running_sum_matches_oracle(ctx, k, n_inputs):
    n = 1 << k
    f = [O.fill_fr(n, 40 + j + k) for j in range(n_inputs)]
    t = O.fill_fr(n, 70 + k)
    m = O.frs_from_ints([(i * 7) % 5 for i in range(n)])
    beta, p0 = O.fill_fr(2, 99 + k)
    if n > 2:
        f[0][1] = O.fr_from_int((-O.fr_to_int(beta)) % R_MOD)  # zero denominator: inverts to zero
    want = O.logup_running_sum(f, t, m, beta, k, p0)
    out = dev(np.zeros((n, 4), np.uint64))
    ctx.logup_running_sum([dev(c) for c in f], dev(t), dev(m), beta, k, p0, out)
    assert np.array_equal(host(out), want)


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 4095, 4096, 4097, 100000, (1 << 18) + 3, 1 << 20])
@pytest.mark.parametrize("op", [0, 1])
def test_prefix_scan_matches_oracle(ctx, n, op):
    a = O.fill_fr(n, 1234 + n)
    init = O.fill_fr(1, 4321)[0]
    want = O.prefix_scan(op, a, init)
    d = dev(a)
    out = ctx.prefix_scan(op, d, init)
    assert np.array_equal(host(out), want)
    ctx.prefix_scan(op, d, init, out=d)  # in place
    assert np.array_equal(host(d), want)


def test_prefix_product_full_size_telescopes(ctx):
    """2^24 elements (the degree-24 layer's z column): scanning a then scanning a^-1 from the result returns to init --
    a size-independent check at full size, plus spot rows against the oracle on a prefix."""
    import torch

    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randint(-(2**63), 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    a[:, 3] &= 0x0FFFFFFFFFFFFFFF
    init = O.fill_fr(1, 8)[0]
    z = ctx.prefix_scan(0, a, init)
    head = 1 << 12
    assert np.array_equal(host(z[:head]), O.prefix_scan(0, host(a[:head]), init))
    inv = a.clone()
    ctx.batch_invert(inv)
    # total = z[n-1] * a[n-1]; scanning the inverses from `total` must land on init * a[n-1]^-1 ... check the closed form
    last = host(z[n - 1:n])[0]
    total = O.fr_mul(last, host(a[n - 1:n])[0])
    back = ctx.prefix_scan(0, inv, total)
    # back[i] = total * prod_{j<i} a_j^-1  =>  back[n-1] * a[n-1]^-1 = init
    fin = O.fr_mul(host(back[n - 1:n])[0], host(inv[n - 1:n])[0])
    assert np.array_equal(fin, init)


@pytest.mark.parametrize("count,n", [(1, 1), (3, 1000), (8, 1 << 14), (0, 64)])
def test_poly_lincomb_matches_bigint(ctx, count, n):
    """out = sum_j s_j p_j in one pass (the SHPLONK prover's per-rotation-set combination), also with out aliasing p_0"""
    polys = [O.fill_fr(n, 800 + j) for j in range(count)]
    sc = O.fill_fr(max(count, 1), 900)[:count]
    pi = [O.frs_to_ints(p) for p in polys]
    si = O.frs_to_ints(sc)
    want = [sum(s * p[i] for s, p in zip(si, pi)) % R_MOD for i in range(n)]
    d = [dev(p) for p in polys]
    out = dev(np.zeros((n, 4), np.uint64))
    ctx.poly_lincomb(d, sc, out)
    assert O.frs_to_ints(host(out)) == want
    if count:
        ctx.poly_lincomb(d, sc, d[0])
        assert O.frs_to_ints(host(d[0])) == want


@pytest.mark.parametrize("n", [1, 63, 2047, 2048, 2049, 4097, 131071, 131072 + 64, (1 << 20) + 5, 1 << 22])
def test_batch_invert_levels_and_zeros(ctx, n):
    """hierarchical Montgomery trick: sizes around the leaf (2048) and slice (64) boundaries, zeros (isolated, a whole strided
    slice, the first and last element) stay zero; checked element-wise against the oracle, larger sizes by a * a^-1 = 1"""
    a = O.fill_fr(n, 5000 + n)
    z = np.zeros(4, np.uint64)
    if n > 1:
        a[0] = z
        a[n - 1] = z
        a[n // 2] = z
    if n > 4096:
        T = (n + 63) // 64
        a[7::T] = z  # every element of one thread's strided slice
    d = dev(a)
    ctx.batch_invert(d)
    got = host(d)
    zero_rows = ~a.any(axis=1)
    assert not got[zero_rows].any()
    if n <= 4097:
        assert np.array_equal(got, O.fr_batch_invert(a))
    else:
        prod = host(ctx.poly_mul(d, dev(a)))
        one = O.fr_from_int(1)
        assert np.array_equal(prod[~zero_rows], np.broadcast_to(one, (int((~zero_rows).sum()), 4)))
        idx = np.arange(0, n, max(1, n // 64))
        assert np.array_equal(got[idx], O.fr_batch_invert(a[idx]))
