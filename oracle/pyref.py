"""ORACLE — TEST INFRASTRUCTURE ONLY.

Independent pure-Python big-integer model of the same mathematics (no shared code
with the C oracle): canonical integers mod r / q, affine BN254 G1 arithmetic,
naive O(n^2) DFT, double-and-add MSM.  Used to pin the C restatement where the
reference ships no direct MSM/NTT input->output vectors.
Small sizes only.
"""
R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
S = 28
GENERATOR = 7
ROOT_OF_UNITY = pow(GENERATOR, (R_MOD - 1) >> S, R_MOD)
ZETA = pow(GENERATOR, 2 * (R_MOD - 1) // 3, R_MOD)  # halo2curves Fr::ZETA


def omega_for(log_n: int) -> int:
    return pow(ROOT_OF_UNITY, 1 << (S - log_n), R_MOD)


def dft(a, omega):
    n = len(a)
    return [sum(a[i] * pow(omega, i * j, R_MOD) for i in range(n)) % R_MOD for j in range(n)]


# ------------------------------------------------------------------ G1 (affine, None = identity)
def g1_add(P, Q):
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    return (x3, (lam * (x1 - x3) - y1) % Q_MOD)


def g1_neg(P):
    return None if P is None else (P[0], (-P[1]) % Q_MOD)


def g1_mul(P, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


G1_GEN = (1, 2)


def msm(scalars, points):
    acc = None
    for s, P in zip(scalars, points):
        acc = g1_add(acc, g1_mul(P, s))
    return acc


def on_curve(P):
    return P is None or (P[1] * P[1] - P[0] ** 3 - 3) % Q_MOD == 0


def compress(P) -> bytes:
    if P is None:
        return bytes(31) + b"\x80"
    v = P[0] | ((P[1] & 1) << 254)
    return v.to_bytes(32, "little")


# ------------------------------------------------------------------ EvaluationDomain semantics
def lagrange_to_coeff(a, k):
    n = 1 << k
    w_inv = pow(omega_for(k), -1, R_MOD)
    n_inv = pow(n, -1, R_MOD)
    return [x * n_inv % R_MOD for x in dft(a, w_inv)]


def coeff_to_extended(a, k, ext_k):
    b = [x * pow(ZETA, i, R_MOD) % R_MOD for i, x in enumerate(a)] + [0] * ((1 << ext_k) - len(a))
    return dft(b, omega_for(ext_k))


def extended_to_coeff(a, ext_k):
    n = 1 << ext_k
    w_inv = pow(omega_for(ext_k), -1, R_MOD)
    n_inv = pow(n, -1, R_MOD)
    zi = pow(ZETA, -1, R_MOD)
    return [x * n_inv % R_MOD * pow(zi, i, R_MOD) % R_MOD for i, x in enumerate(dft(a, w_inv))]


def eval_poly(p, x):
    acc = 0
    for c in reversed(p):
        acc = (acc * x + c) % R_MOD
    return acc
