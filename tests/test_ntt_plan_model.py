"""Executable model of the device NTT's pass decomposition (csrc/ntt.cu), in pure Python big ints.

Mirrors the kernel's index algebra one to one -- digits n_1..n_P (most significant first), in-place strided tiles
with `base`, twist c = digit-reversal of the already transformed digits, global twiddle w_{2^(t+s)}^(K 2^t + c),
transposed store of the last pass, zero padding + zeta^i pre-scaling on load, scaled / coset post-store -- and checks
it against the naive DFT for every digit split, including splits the device only uses at 2^25..2^28 (4 passes), by
running the same algebra with a small maximum digit.  Needs no GPU: it pins the ALGORITHM; the -m gpu parity tests
pin the CUDA implementation of it.
"""
import random

import pytest

from oracle import pyref as P

R = P.R_MOD
C = 8  # lanes per tile in the kernel; the model iterates lanes explicitly


def plan_digits(log_n, max_digit):
    if log_n <= max_digit:
        return [log_n]
    p = -(-log_n // max_digit)
    base, extra = divmod(log_n, p)
    return [base + (1 if i < extra else 0) for i in range(p)]


def brev(x, bits):
    return int(bin(x)[2:].zfill(bits)[::-1], 2) if bits else 0


def model_ntt(a_in, log_in, log_n, omega, digits, pre=False, post=None):
    """post: None or list of 3 multipliers indexed by output index % 3 (scale * zeta^-i pattern)."""
    n = log_n
    N = 1 << n
    tab = {}  # universal stage table: (u, j) -> w_{2^u}^j

    def tw(u, j):
        if (u, j) not in tab:
            tab[(u, j)] = pow(pow(omega, 1 << (n - u), R), j, R)
        return tab[(u, j)]

    zeta = P.ZETA
    work = [0] * N
    out = [0] * N
    Pn = len(digits)
    t = 0
    for p, m in enumerate(digits):
        L = 1 << m
        rest = n - t - m
        last = p == Pn - 1
        src = a_in if p == 0 else work
        if not last:
            for o in range(1 << t):
                # c = digit reversal of o (o = [K_1][K_2].. positionally, K_1 most significant)
                sh, tq, c = t, 0, 0
                for q in range(p):
                    sh -= digits[q]
                    c |= ((o >> sh) & ((1 << digits[q]) - 1)) << tq
                    tq += digits[q]
                for r in range(1 << rest):
                    base = (o << (n - t)) + r
                    tile = [0] * L
                    for d in range(L):
                        gi = base + (d << rest)
                        v = 0
                        if p > 0 or gi < (1 << log_in):
                            v = src[gi]
                            if p == 0 and pre:
                                v = v * pow(zeta, gi % 3, R) % R
                        tile[brev(d, m)] = v
                    for s in range(1, m + 1):
                        half = 1 << (s - 1)
                        for bb in range(L // 2):
                            K, blk = bb & (half - 1), bb >> (s - 1)
                            p0 = (blk << s) + K
                            w = tw(t + s, (K << t) + c)
                            u, v = tile[p0], tile[p0 + half] * w % R
                            tile[p0], tile[p0 + half] = (u + v) % R, (u - v) % R
                    for K in range(L):
                        work[base + (K << rest)] = tile[K]
        else:
            for c in range(1 << t):
                # row position of twist c: K_q sits at bit offset n - t_q - n_q
                tq, pos = 0, 0
                for q in range(Pn - 1):
                    kq = (c >> tq) & ((1 << digits[q]) - 1)
                    tq += digits[q]
                    pos |= kq << (n - tq)
                tile = [0] * L
                for d in range(L):
                    gi = pos + d
                    v = 0
                    if p > 0 or gi < (1 << log_in):
                        v = src[gi]
                        if p == 0 and pre:
                            v = v * pow(zeta, gi % 3, R) % R
                    tile[brev(d, m)] = v
                for s in range(1, m + 1):
                    half = 1 << (s - 1)
                    for bb in range(L // 2):
                        K, blk = bb & (half - 1), bb >> (s - 1)
                        p0 = (blk << s) + K
                        w = tw(t + s, (K << t) + c)
                        u, v = tile[p0], tile[p0 + half] * w % R
                        tile[p0], tile[p0 + half] = (u + v) % R, (u - v) % R
                for K in range(L):
                    go = (K << t) + c
                    v = tile[K]
                    if post is not None:
                        v = v * post[go % 3] % R
                    out[go] = v
        t += m
    return out


@pytest.mark.parametrize("log_n,max_digit", [(1, 8), (3, 8), (6, 8), (4, 2), (5, 2), (6, 2), (7, 3), (8, 2), (9, 3), (10, 4)])
def test_pass_decomposition_is_the_dft(log_n, max_digit):
    rng = random.Random(log_n * 31 + max_digit)
    a = [rng.randrange(R) for _ in range(1 << log_n)]
    w = P.omega_for(log_n)
    digits = plan_digits(log_n, max_digit)
    assert sum(digits) == log_n and max(digits) <= max_digit
    assert model_ntt(a, log_n, log_n, w, digits) == P.dft(a, w)


@pytest.mark.parametrize("k,max_digit", [(3, 2), (4, 3), (5, 2)])
def test_fused_domain_transforms(k, max_digit):
    """coeff_to_extended (zero padding + zeta^i on load) and extended_to_coeff (n^-1 zeta^-i on store) in the model."""
    rng = random.Random(k)
    ek = k + 2
    coeff = [rng.randrange(R) for _ in range(1 << k)]
    digits = plan_digits(ek, max_digit)
    ext = model_ntt(coeff, k, ek, P.omega_for(ek), digits, pre=True)
    assert ext == P.coeff_to_extended(coeff, k, ek)
    ninv = pow(1 << ek, -1, R)
    zi = pow(P.ZETA, -1, R)
    post = [ninv, ninv * zi % R, ninv * zi * zi % R]  # n^-1 * zeta^-(i mod 3)  (zeta^3 = 1)
    back = model_ntt(ext, ek, ek, pow(P.omega_for(ek), -1, R), digits, post=post)
    assert back == coeff + [0] * ((1 << ek) - (1 << k))
