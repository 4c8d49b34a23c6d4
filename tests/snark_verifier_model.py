"""TEST INFRASTRUCTURE: a big-integer model of snark-verifier's native PLONK verifier for the reference's Poseidon-transcript
proofs (the chunk proofs of integration/tests/test_data/full_proof_*.json, protocol = release-v0.13.1/chunk.protocol):

  * Poseidon sponge (snark-verifier util/hash/poseidon.rs) with T = 5, RATE = 4, R_F = 8, R_P = 60 -- the parameter set under
    which the reference's shipped proofs verify (T = 3 / R_P = 57 and others were tried and rejected by the pairing) -- and
    constants from the Grain LFSR of the Poseidon paper (generate_parameters_grain; the t = 3 output reproduces the well-known
    first round constant 0x0ee9a592...cd8e6e), Cauchy MDS from the same stream;
  * PoseidonTranscript (system/halo2/transcript/halo2.rs): scalars absorbed as is, points as (x mod r, y mod r), challenges =
    squeeze; proof = compressed little-endian points and little-endian scalars;
  * PlonkProof::read / PlonkVerifier::verify (verifier/plonk.rs) driven by the protocol JSON: witness phases, quotient chunks,
    evaluations, the quotient numerator expression tree, the opening queries;
  * Bdfg21 (pcs/kzg/multiopen/bdfg21.rs): query sets in first-appearance order, powers of mu inside a set, powers of gamma across
    sets, normalisation by the first set; result = a KZG accumulator (lhs, rhs) decided with one pairing.
Upstream pins: snark-verifier @ 948671c (/root/reference/Cargo.lock:3948-3950).  Everything here is plain Python integers.
"""
from __future__ import annotations

from pairing_model import Q, R, g1_add, g1_mul

P = R  # the scalar field (Fr): Poseidon's field


# ------------------------------------------------------------------------------------------------ Grain LFSR -> Poseidon constants
class Grain:
    def __init__(self, t: int, r_f: int, r_p: int, n_bits: int = 254, field: int = 1, sbox: int = 0):
        bits = []
        for value, width in ((field, 2), (sbox, 4), (n_bits, 12), (t, 12), (r_f, 10), (r_p, 10), ((1 << 30) - 1, 30)):
            bits += [(value >> (width - 1 - i)) & 1 for i in range(width)]
        assert len(bits) == 80
        self.s = bits
        for _ in range(160):
            self._update()

    def _update(self) -> int:
        s = self.s
        b = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0]
        s.pop(0)
        s.append(b)
        return b

    def bit(self) -> int:  # self-shrinking: (1, b) -> b ; (0, _) -> nothing
        while True:
            first = self._update()
            second = self._update()
            if first:
                return second

    def bits_int(self, n: int) -> int:
        v = 0
        for _ in range(n):
            v = (v << 1) | self.bit()
        return v

    def field_element(self, n_bits: int = 254) -> int:  # with rejection
        while True:
            v = self.bits_int(n_bits)
            if v < P:
                return v

    def field_element_no_rejection(self, n_bits: int = 254) -> int:
        return self.bits_int(n_bits) % P


class PoseidonSpec:
    def __init__(self, t: int = 5, r_f: int = 8, r_p: int = 60, secure_mds: int = 0):
        g = Grain(t, r_f, r_p)
        self.t, self.r_f, self.r_p = t, r_f, r_p
        self.rc = [[g.field_element() for _ in range(t)] for _ in range(r_f + r_p)]
        select = secure_mds
        while True:
            vals = [g.field_element_no_rejection() for _ in range(2 * t)]
            if len(set(vals)) != 2 * t:
                continue
            if select:
                select -= 1
                continue
            xs, ys = vals[:t], vals[t:]
            break
        self.mds = [[pow((xs[i] + ys[j]) % P, -1, P) for j in range(t)] for i in range(t)]

    def permute(self, state):
        t, half = self.t, self.r_f // 2
        rnd = 0

        def mix(s):
            return [sum(self.mds[i][j] * s[j] for j in range(t)) % P for i in range(t)]

        for _ in range(half):
            state = mix([pow((x + c) % P, 5, P) for x, c in zip(state, self.rc[rnd])])
            rnd += 1
        for _ in range(self.r_p):
            state = [(x + c) % P for x, c in zip(state, self.rc[rnd])]
            state[0] = pow(state[0], 5, P)
            state = mix(state)
            rnd += 1
        for _ in range(half):
            state = mix([pow((x + c) % P, 5, P) for x, c in zip(state, self.rc[rnd])])
            rnd += 1
        return state


class Poseidon:
    """snark-verifier's sponge: state = [2^64, 0, 0]; update() buffers; squeeze() absorbs the buffer in RATE chunks (a partial or
    empty last chunk is followed by a 1), permutes, and returns state[1]; the state carries over between squeezes."""

    def __init__(self, spec: PoseidonSpec, rate: int = 0):
        self.spec, self.rate = spec, rate or spec.t - 1
        self.state = [1 << 64] + [0] * (spec.t - 1)
        self.buf = []

    def update(self, elems):
        self.buf += [e % P for e in elems]

    def _permutation(self, chunk):
        s = self.state[:]
        for i, v in enumerate(chunk):
            s[i + 1] = (s[i + 1] + v) % P
        if len(chunk) + 1 < self.spec.t:
            s[len(chunk) + 1] = (s[len(chunk) + 1] + 1) % P
        self.state = self.spec.permute(s)

    def squeeze(self) -> int:
        buf, self.buf = self.buf, []
        exact = len(buf) % self.rate == 0
        for i in range(0, len(buf), self.rate):
            self._permutation(buf[i:i + self.rate])
        if exact:
            self._permutation([])
        return self.state[1]


# ------------------------------------------------------------------------------------------------ transcript over a proof
def decompress_g1(b: bytes):
    """halo2curves compressed G1: x little-endian, bit 254 = lsb(y) ("sign"), bit 255 = identity"""
    v = int.from_bytes(b, "little")
    if v >> 255:
        return None
    sign = (v >> 254) & 1
    x = v & ((1 << 254) - 1)
    y2 = (x * x * x + 3) % Q
    y = pow(y2, (Q + 1) // 4, Q)
    if y * y % Q != y2:
        raise ValueError("not on the curve")
    if (y & 1) != sign:
        y = Q - y
    return (x, y)


class PoseidonTranscript:
    def __init__(self, proof: bytes, spec: PoseidonSpec):
        self.h, self.proof, self.pos = Poseidon(spec), proof, 0

    def common_scalar(self, s):
        self.h.update([s])

    def common_point(self, p):
        self.h.update([p[0] % P, p[1] % P])  # fe_to_fe::<Fq, Fr>: the coordinates reduced into the scalar field

    def read_point(self):
        p = decompress_g1(self.proof[self.pos:self.pos + 32])
        self.pos += 32
        self.common_point(p)
        return p

    def read_scalar(self):
        s = int.from_bytes(self.proof[self.pos:self.pos + 32], "little")
        if s >= P:
            raise ValueError("scalar not canonical")
        self.pos += 32
        self.common_scalar(s)
        return s

    def squeeze(self):
        return self.h.squeeze()


# ------------------------------------------------------------------------------------------------ the PLONK verifier
R_MONT_INV = pow(1 << 256, -1, P)


def limbs_to_int(l4) -> int:  # the protocol stores field elements as raw Montgomery limbs
    return sum(int(v) << (64 * i) for i, v in enumerate(l4)) * R_MONT_INV % P


def limbs_to_fq(l4) -> int:
    return sum(int(v) << (64 * i) for i, v in enumerate(l4)) * pow(1 << 256, -1, Q) % Q


def msm(pairs):
    acc = None
    for s, p in pairs:
        s %= P
        if s and p is not None:
            acc = g1_add(acc, g1_mul(p, s))
    return acc


def verify_plonk(protocol: dict, instances, proof: bytes, spec: PoseidonSpec):
    """returns (lhs, rhs): the KZG accumulator of the proof -- valid iff e(lhs, g2) = e(rhs, s_g2)"""
    k, n = protocol["domain"]["k"], protocol["domain"]["n"]
    omega = limbs_to_int(protocol["domain"]["gen"])
    omega_inv = limbs_to_int(protocol["domain"]["gen_inv"])
    n_inv = limbs_to_int(protocol["domain"]["n_inv"])
    assert pow(omega, n, P) == 1 and omega * omega_inv % P == 1 and n * n_inv % P == 1
    preprocessed = [(limbs_to_fq(p["x"]), limbs_to_fq(p["y"])) for p in protocol["preprocessed"]]
    tr = PoseidonTranscript(proof, spec)
    if protocol["transcript_initial_state"] is not None:
        tr.common_scalar(limbs_to_int(protocol["transcript_initial_state"]))
    assert protocol["instance_committing_key"] is None and [len(i) for i in instances] == protocol["num_instance"]
    for col in instances:
        for v in col:
            tr.common_scalar(v)
    witnesses, challenges = [], []
    for nw, nc in zip(protocol["num_witness"], protocol["num_challenge"]):
        witnesses += [tr.read_point() for _ in range(nw)]
        challenges += [tr.squeeze() for _ in range(nc)]
    quotients = [tr.read_point() for _ in range(protocol["quotient"]["num_chunk"])]
    z = tr.squeeze()
    evals = {(e["poly"], e["rotation"]): tr.read_scalar() for e in protocol["evaluations"]}
    mu, gamma = tr.squeeze(), tr.squeeze()
    w = tr.read_point()
    z_prime = tr.squeeze()
    w_prime = tr.read_point()
    assert tr.pos == len(proof)

    zn = pow(z, n, P)

    def lagrange(i):
        wi = pow(omega, i, P) if i >= 0 else pow(omega_inv, -i, P)
        return wi * n_inv % P * (zn - 1) % P * pow((z - wi) % P, -1, P) % P

    n_pre, n_inst = len(preprocessed), len(instances)
    quotient_poly = n_pre + n_inst + len(witnesses)

    def poly_eval(poly, rot):
        if (poly, rot) in evals:
            return evals[(poly, rot)]
        if n_pre <= poly < n_pre + n_inst:  # instance column not committed: interpolate
            col = instances[poly - n_pre]
            return sum(v * lagrange(i - rot) for i, v in enumerate(col) if v) % P
        raise KeyError((poly, rot))

    def ev(e):
        (kind, arg), = e.items()
        if kind == "Constant":
            return limbs_to_int(arg)
        if kind == "CommonPolynomial":
            if arg == "Identity":
                return z
            return lagrange(arg["Lagrange"])
        if kind == "Polynomial":
            return poly_eval(arg["poly"], arg["rotation"])
        if kind == "Challenge":
            return challenges[arg]
        if kind == "Negated":
            return -ev(arg) % P
        if kind == "Sum":
            return (ev(arg[0]) + ev(arg[1])) % P
        if kind == "Product":
            return ev(arg[0]) * ev(arg[1]) % P
        if kind == "Scaled":
            return ev(arg[0]) * limbs_to_int(arg[1]) % P
        if kind == "DistributePowers":
            exprs, base = arg
            b = ev(base)
            acc = ev(exprs[0])
            for x in exprs[1:]:
                acc = (acc * b + ev(x)) % P
            return acc
        raise NotImplementedError(kind)

    quotient_eval = ev(protocol["quotient"]["numerator"]) * pow((zn - 1) % P, -1, P) % P
    evals[(quotient_poly, 0)] = quotient_eval
    commitments = preprocessed + [None] * n_inst + witnesses
    zn_c = pow(zn, protocol["quotient"]["chunk_degree"], P)
    quotient_terms = [(pow(zn_c, i, P), q) for i, q in enumerate(quotients)]
    commitments.append(msm(quotient_terms))

    # ---- Bdfg21
    polys_order, per_poly = [], {}
    for q in protocol["queries"]:
        shift = pow(omega, q["rotation"], P) if q["rotation"] >= 0 else pow(omega_inv, -q["rotation"], P)
        if q["poly"] not in per_poly:
            per_poly[q["poly"]] = ([], [])
            polys_order.append(q["poly"])
        sh, evs = per_poly[q["poly"]]
        if shift not in sh:
            sh.append(shift)
            evs.append(evals[(q["poly"], q["rotation"])])
    sets = []  # [shifts, polys, evals-per-poly]
    for poly in polys_order:
        sh, evs = per_poly[poly]
        for s in sets:
            if set(s[0]) == set(sh):
                if poly not in s[1]:
                    s[1].append(poly)
                    s[2].append([evs[sh.index(x)] for x in s[0]])
                break
        else:
            sets.append([sh, [poly], [evs]])
    z_s = []
    for sh, _, _ in sets:
        v = 1
        for s in sh:
            v = v * (z_prime - s * z) % P
        z_s.append(v)
    terms, constant = [], 0
    for i, (sh, polys, evs) in enumerate(sets):
        coeff = pow(gamma, i, P) * z_s[0] % P * pow(z_s[i], -1, P) % P
        pts = [s * z % P for s in sh]
        for j, (poly, ev_list) in enumerate(zip(polys, evs)):
            # r(z') by Lagrange interpolation through (pts, ev_list)
            r_eval = 0
            for a, (xa, ya) in enumerate(zip(pts, ev_list)):
                num = den = 1
                for b_, xb in enumerate(pts):
                    if b_ != a:
                        num = num * (z_prime - xb) % P
                        den = den * (xa - xb) % P
                r_eval = (r_eval + ya * num % P * pow(den, -1, P)) % P
            c = coeff * pow(mu, j, P) % P
            terms.append((c, commitments[poly]))
            constant = (constant + c * r_eval) % P
    terms.append((-constant % P, (1, 2)))
    terms.append((-z_s[0] % P, w))
    f = msm(terms)
    lhs = g1_add(f, g1_mul(w_prime, z_prime)) if z_prime else f
    # msm_terms / quotient_terms: the two multi-scalar multiplications of this verification as (scalar, point) lists, so that a test
    # can redo them with another MSM implementation (f = msm(msm_terms) is the opening's left-hand side before + z' W')
    return lhs, w_prime, {"challenges": challenges, "z": z, "mu": mu, "gamma": gamma, "z_prime": z_prime, "n_sets": len(sets),
                          "msm_terms": terms, "quotient_terms": quotient_terms, "f": f}
