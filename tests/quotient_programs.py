"""Shared helpers of the quotient-construction tests: random GraphEvaluator programs and an independent big-integer model.

A program is (calcs, constants, rotations): calcs = [(op, a, b, parts)], sources = (kind, index, rotation_index); the same
neutral form is packed for the oracle (oracle.pack_program), the host emulation and the CUDA path (zk.Graph).
"""
import random

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ZETA = pow(7, 2 * (R_MOD - 1) // 3, R_MOD)
ROOT_OF_UNITY = pow(7, (R_MOD - 1) >> 28, R_MOD)
DELTA = pow(7, 1 << 28, R_MOD)

(S_CONST, S_INTER, S_FIXED, S_ADVICE, S_INSTANCE, S_CHALLENGE, S_BETA, S_GAMMA, S_THETA, S_Y, S_PREV, S_X) = range(12)
(C_ADD, C_SUB, C_MUL, C_SQUARE, C_DOUBLE, C_NEGATE, C_HORNER, C_STORE) = range(8)


def omega_of(log_n: int) -> int:
    return pow(ROOT_OF_UNITY, 1 << (28 - log_n), R_MOD)


def random_program(seed: int, n_calcs: int, n_fixed: int, n_advice: int, n_instance: int, n_challenges: int, n_rot: int,
                   n_constants: int = 4, use_prev: bool = True, use_x: bool = True, chain_bias: float = 0.5):
    """Random program; chain_bias = probability that an operand is a recent intermediate (controls live ranges)."""
    rng = random.Random(seed)
    rotations = [0] + rng.sample(range(-5, 6), k=min(n_rot - 1, 10)) if n_rot > 1 else [0]
    rotations = rotations[:n_rot]
    constants = [rng.randrange(R_MOD) for _ in range(n_constants)]
    if n_constants > 1:
        constants[0], constants[1] = 0, 1

    def source(i):
        if i > 0 and rng.random() < chain_bias:
            lo = max(0, i - 6) if rng.random() < 0.8 else 0
            return (S_INTER, rng.randrange(lo, i), 0)
        kinds = [S_CONST, S_BETA, S_GAMMA, S_THETA, S_Y]
        if n_fixed: kinds += [S_FIXED] * 2
        if n_advice: kinds += [S_ADVICE] * 4
        if n_instance: kinds += [S_INSTANCE]
        if n_challenges: kinds += [S_CHALLENGE]
        if use_prev: kinds += [S_PREV]
        if use_x: kinds += [S_X]
        k = rng.choice(kinds)
        if k == S_CONST: return (k, rng.randrange(n_constants), 0)
        if k == S_FIXED: return (k, rng.randrange(n_fixed), rng.randrange(len(rotations)))
        if k == S_ADVICE: return (k, rng.randrange(n_advice), rng.randrange(len(rotations)))
        if k == S_INSTANCE: return (k, rng.randrange(n_instance), rng.randrange(len(rotations)))
        if k == S_CHALLENGE: return (k, rng.randrange(n_challenges), 0)
        return (k, 0, 0)

    calcs = []
    for i in range(n_calcs):
        op = rng.choice([C_ADD, C_SUB, C_MUL, C_MUL, C_MUL, C_SQUARE, C_DOUBLE, C_NEGATE, C_HORNER, C_STORE])
        if op in (C_ADD, C_SUB, C_MUL):
            calcs.append((op, source(i), source(i), None))
        elif op == C_HORNER:
            calcs.append((op, source(i), source(i), [source(i) for _ in range(rng.randrange(0, 6))]))
        else:
            calcs.append((op, source(i), None, None))
    return calcs, constants, rotations


def model_evaluate(calcs, constants, rotations, fixed, advice, instance, challenges, beta, gamma, theta, y, ext_omega, values,
                   log_size: int, rot_scale: int):
    """GraphEvaluator::evaluate over all rows with Python integers (canonical values, not Montgomery)."""
    size = 1 << log_size
    out = []
    for idx in range(size):
        rows = [(idx + r * rot_scale) % size for r in rotations]
        inter = []
        x = ZETA * pow(ext_omega, idx, R_MOD) % R_MOD if ext_omega is not None else ZETA

        def get(s):
            k, i, r = s
            if k == S_CONST: return constants[i]
            if k == S_INTER: return inter[i]
            if k == S_FIXED: return fixed[i][rows[r]]
            if k == S_ADVICE: return advice[i][rows[r]]
            if k == S_INSTANCE: return instance[i][rows[r]]
            if k == S_CHALLENGE: return challenges[i]
            if k == S_BETA: return beta
            if k == S_GAMMA: return gamma
            if k == S_THETA: return theta
            if k == S_Y: return y
            if k == S_PREV: return values[idx]
            return x

        for op, a, b, parts in calcs:
            if op == C_ADD: v = get(a) + get(b)
            elif op == C_SUB: v = get(a) - get(b)
            elif op == C_MUL: v = get(a) * get(b)
            elif op == C_SQUARE: v = get(a) ** 2
            elif op == C_DOUBLE: v = 2 * get(a)
            elif op == C_NEGATE: v = -get(a)
            elif op == C_HORNER:
                f = get(b)
                v = get(a)
                for p in parts:
                    v = (v * f + get(p)) % R_MOD
            else: v = get(a)
            inter.append(v % R_MOD)
        out.append(inter[-1] if inter else 0)
    return out


def model_permutation_product(values, sigma, beta, gamma, delta_omega_start, delta, omega, k, z_init):
    n = 1 << k
    mv = [1] * n
    for v, s in zip(values, sigma):
        for i in range(n):
            mv[i] = mv[i] * (beta * s[i] + gamma + v[i]) % R_MOD
    mv = [pow(t, -1, R_MOD) if t else 0 for t in mv]
    d0 = delta_omega_start
    for v in values:
        dw = d0
        for i in range(n):
            mv[i] = mv[i] * (dw * beta + gamma + v[i]) % R_MOD
            dw = dw * omega % R_MOD
        d0 = d0 * delta % R_MOD
    z = [z_init % R_MOD]
    for i in range(1, n):
        z.append(z[-1] * mv[i - 1] % R_MOD)
    return z


def model_logup(inputs, table, m, beta, k, phi_init):
    n = 1 << k
    inv = lambda t: pow(t, -1, R_MOD) if t % R_MOD else 0
    d = [(sum(inv(f[i] + beta) for f in inputs) - m[i] * inv(table[i] + beta)) % R_MOD for i in range(n)]
    phi = [phi_init % R_MOD]
    for i in range(1, n):
        phi.append((phi[-1] + d[i - 1]) % R_MOD)
    return phi
