"""Independent big-integer model of the BN254 optimal-ate pairing (test infrastructure only).

Used for the verifier-side known-answer test SURVEY.md §8(c) #3 / §8(f).4 names: the KZG accumulators the reference
ships (release-v0.13.1/proof.data, integration/tests/test_data/full_proof_1.json) must satisfy the pairing equation of
release-v0.13.1/evm_verifier.yul:1230-1240 (EIP-197 precompile 0x08: prod e(P_i, Q_i) == 1).

Construction (textbook, unoptimised): Fq12 = Fq[w] / (w^12 - 18 w^6 + 82); the Fq2 element a + b*i embeds as
a - 9b + b*w^6 (i = w^6 - 9); G2 is the D-type sextic twist y^2 = x^3 + 3/(9 + i), untwisted by (x w^2, y w^3);
Miller loop over 6u + 2 = 29793968203157093288 plus the two Frobenius steps; final exponentiation by (q^12 - 1)/r.
"""
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ATE_LOOP_COUNT = 29793968203157093288
LOG_ATE = 63
# w^12 = 18 w^6 - 82
_MOD_HI, _MOD_LO = 18, -82


class F12:
    __slots__ = ("c",)

    def __init__(self, c):
        self.c = [x % Q for x in c]

    @staticmethod
    def one():
        return F12([1] + [0] * 11)

    @staticmethod
    def zero():
        return F12([0] * 12)

    def __add__(self, o):
        return F12([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return F12([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return F12([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k):
        return F12([a * k for a in self.c])

    def __mul__(self, o):
        t = [0] * 23
        for i, a in enumerate(self.c):
            if a:
                for j, b in enumerate(o.c):
                    t[i + j] += a * b
        for k in range(22, 11, -1):  # w^k = 18 w^(k-6) - 82 w^(k-12)
            v = t[k]
            if v:
                t[k - 6] += _MOD_HI * v
                t[k - 12] += _MOD_LO * v
        return F12(t[:12])

    def __pow__(self, e):
        acc, base = F12.one(), self
        while e:
            if e & 1:
                acc = acc * base
            base = base * base
            e >>= 1
        return acc

    def inv(self):
        """solve (self * x = 1) as a 12 x 12 linear system over Fq: column j of the matrix is self * w^j"""
        cols = []
        v = self
        for _ in range(12):
            cols.append(v.c)
            v = v * W
        m = [[cols[j][i] for j in range(12)] + [1 if i == 0 else 0] for i in range(12)]
        for col in range(12):
            piv = next(r for r in range(col, 12) if m[r][col])
            m[col], m[piv] = m[piv], m[col]
            inv = pow(m[col][col], -1, Q)
            m[col] = [x * inv % Q for x in m[col]]
            for r in range(12):
                if r != col and m[r][col]:
                    f = m[r][col]
                    m[r] = [(x - f * y) % Q for x, y in zip(m[r], m[col])]
        return F12([m[i][12] for i in range(12)])


W = F12([0, 1] + [0] * 10)


def embed_fq(a):
    return F12([a] + [0] * 11)


def embed_fq2(c0, c1):
    """a = c0 + c1*i, i = w^6 - 9"""
    return F12([c0 - 9 * c1] + [0] * 5 + [c1] + [0] * 5)


def twist(q2):
    """G2 affine ((x0, x1), (y0, y1)) -> point of E(Fq12): (x w^2, y w^3)"""
    (x0, x1), (y0, y1) = q2
    return embed_fq2(x0, x1) * (W ** 2), embed_fq2(y0, y1) * (W ** 3)


def cast_g1(p):
    return embed_fq(p[0]), embed_fq(p[1])


def _div(a, b):
    return a * b.inv()


def _double(p):
    x, y = p
    m = _div((x * x).scale(3), y.scale(2))
    nx = m * m - x.scale(2)
    return nx, m * (x - nx) - y


def _add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        return _double(p1) if y1 == y2 else None
    m = _div(y2 - y1, x2 - x1)
    nx = m * m - x1 - x2
    return nx, m * (x1 - nx) - y1


def _line(p1, p2, t):
    """the line through p1, p2 (tangent if equal) evaluated at t"""
    x1, y1 = p1
    x2, y2 = p2
    xt, yt = t
    if not x1 == x2:
        m = _div(y2 - y1, x2 - x1)
        return m * (xt - x1) - (yt - y1)
    if y1 == y2:
        m = _div((x1 * x1).scale(3), y1.scale(2))
        return m * (xt - x1) - (yt - y1)
    return xt - x1


def miller_loop(q2, p1):
    """q2: G2 affine over Fq2 (or None), p1: G1 affine (or None); the value BEFORE the final exponentiation"""
    if q2 is None or p1 is None:
        return F12.one()
    qq, pp = twist(q2), cast_g1(p1)
    r, f = qq, F12.one()
    for i in range(LOG_ATE, -1, -1):
        f = f * f * _line(r, r, pp)
        r = _double(r)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _line(r, qq, pp)
            r = _add(r, qq)
    q1 = (qq[0] ** Q, qq[1] ** Q)
    nq2 = (q1[0] ** Q, -(q1[1] ** Q))
    f = f * _line(r, q1, pp)
    r = _add(r, q1)
    f = f * _line(r, nq2, pp)
    return f


def final_exponentiation(f):
    return f ** ((Q ** 12 - 1) // R)


def pairing(q2, p1):
    return final_exponentiation(miller_loop(q2, p1))


def pairing_check(pairs) -> bool:
    """EIP-197: prod e(P_i, Q_i) == 1 for pairs (P_i in G1, Q_i in G2)"""
    f = F12.one()
    for p1, q2 in pairs:
        f = f * miller_loop(q2, p1)
    return final_exponentiation(f) == F12.one()


# ---- plain affine arithmetic on G1 / G2 for building test points
def g1_add(p, q):
    if p is None: return q
    if q is None: return p
    if p[0] == q[0]:
        if (p[1] + q[1]) % Q == 0: return None
        m = 3 * p[0] * p[0] * pow(2 * p[1], -1, Q) % Q
    else:
        m = (q[1] - p[1]) * pow(q[0] - p[0], -1, Q) % Q
    x = (m * m - p[0] - q[0]) % Q
    return x, (m * (p[0] - x) - p[1]) % Q


def g1_mul(p, k):
    acc = None
    while k:
        if k & 1: acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def _f2mul(a, b):
    return (a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q


def _f2inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
    return a[0] * d % Q, -a[1] * d % Q


def _f2sub(a, b):
    return (a[0] - b[0]) % Q, (a[1] - b[1]) % Q


def g2_add(p, q):
    if p is None: return q
    if q is None: return p
    if p[0] == q[0]:
        if ((p[1][0] + q[1][0]) % Q, (p[1][1] + q[1][1]) % Q) == (0, 0): return None
        xx = _f2mul(p[0], p[0])
        m = _f2mul((3 * xx[0] % Q, 3 * xx[1] % Q), _f2inv((2 * p[1][0] % Q, 2 * p[1][1] % Q)))
    else:
        m = _f2mul(_f2sub(q[1], p[1]), _f2inv(_f2sub(q[0], p[0])))
    x = _f2sub(_f2sub(_f2mul(m, m), p[0]), q[0])
    return x, _f2sub(_f2mul(m, _f2sub(p[0], x)), p[1])


def g2_mul(p, k):
    acc = None
    while k:
        if k & 1: acc = g2_add(acc, p)
        p = g2_add(p, p)
        k >>= 1
    return acc


def g2_neg(p):
    return p[0], (-p[1][0] % Q, -p[1][1] % Q)


def g2_on_curve(p):
    (x, y) = p
    b2 = _f2mul((3, 0), _f2inv((9, 1)))
    lhs = _f2mul(y, y)
    x3 = _f2mul(_f2mul(x, x), x)
    return lhs == ((x3[0] + b2[0]) % Q, (x3[1] + b2[1]) % Q)


G1_GEN = (1, 2)
# EIP-197 / evm_verifier.yul:1230-1233 word order is (x_c1, x_c0, y_c1, y_c0); here points are ((c0, c1), (c0, c1))
G2_GEN = ((0x1800DEEF121F1E76426A00665E5C4479674322D4F75EDADD46DEBD5CD992F6ED, 0x198E9393920D483A7260BFB731FB5D25F1AA493335A9E71297E485B7AEF312C2),
          (0x12C85EA5DB8C6DEB4AAB71808DCB408FE3D1E7690C43D37B4CE6CC0166FA7DAA, 0x090689D0585FF075EC9E99AD690C3395BC4B313370B38EF355ACDADCD122975B))
