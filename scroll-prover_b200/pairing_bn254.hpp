// pairing_bn254.hpp — host-side BN254 optimal-ate pairing for the verifier-side checks of SURVEY.md §8(f).4.
//
// This is HOST code by design (the reference decides accumulators on the CPU as well: snark-verifier's
// KzgAs / `Bn256::multi_miller_loop(..).final_exponentiation().is_identity()`, reached from
// /root/reference/integration/src/prove.rs:50-53,78-80; on chain it is precompile 0x08 called at
// /root/reference/release-v0.13.1/evm_verifier.yul:1240).  It runs on the field layer of csrc/ff.cuh compiled for the host
// (the same 8 x u32 Montgomery code the device uses, PTX leaves replaced by their 64-bit emulation), so a maintainer can
// check an accumulator, or a KZG opening against [tau]G2, without any other dependency.  One pairing check takes
// milliseconds; nothing here is on the hot path.
//
// Tower: Fq2 = Fq[u]/(u^2 + 1), Fq6 = Fq2[v]/(v^3 - xi), xi = 9 + u, Fq12 = Fq6[w]/(w^2 - v); G2 is the D-type sextic
// twist y^2 = x^3 + 3/xi, untwisted by (x w^2, y w^3).  Miller loop over 6t + 2 with the two Frobenius steps, affine
// line slopes (one Fq2 inversion per step), final exponentiation = easy part by conjugation / inversion / Frobenius^2,
// hard part (q^4 - q^2 + 1)/r by plain square-and-multiply.  Validated in tests/test_pairing_host.py against bilinearity,
// the independent big-integer model (tests/pairing_model.py) and the reference's shipped accumulators.
#pragma once
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

#include "csrc/ff.cuh"

namespace halo2_b200 {
namespace pairing {

using Fq = b200zk::Fq;

struct Fq2 {
    Fq c0, c1;
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
    Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
    Fq2 neg() const { return {c0.neg(), c1.neg()}; }
    Fq2 conj() const { return {c0, c1.neg()}; }
    Fq2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    Fq2 operator*(const Fq2& o) const {  // (a + bu)(c + du) = ac - bd + ((a + b)(c + d) - ac - bd) u
        Fq ac = c0 * o.c0, bd = c1 * o.c1;
        return {ac - bd, (c0 + c1) * (o.c0 + o.c1) - ac - bd};
    }
    Fq2 scale(const Fq& s) const { return {c0 * s, c1 * s}; }
    Fq2 sqr() const { return (*this) * (*this); }
    Fq2 mul_by_xi() const {  // (a + bu)(9 + u) = 9a - b + (a + 9b) u
        Fq a9 = c0.dbl().dbl().dbl() + c0, b9 = c1.dbl().dbl().dbl() + c1;
        return {a9 - c1, c0 + b9};
    }
    Fq2 inv() const {  // 1/(a + bu) = (a - bu)/(a^2 + b^2)
        Fq d = (c0 * c0 + c1 * c1).inv();
        return {c0 * d, (c1 * d).neg()};
    }
    Fq2 pow(const uint64_t* e, int limbs) const {
        Fq2 acc = one();
        for (int i = limbs - 1; i >= 0; --i)
            for (int b = 63; b >= 0; --b) {
                acc = acc.sqr();
                if ((e[i] >> b) & 1) acc = acc * (*this);
            }
        return acc;
    }
};

struct Fq6 {
    Fq2 c0, c1, c2;  // c0 + c1 v + c2 v^2
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    Fq6 operator*(const Fq6& o) const {  // schoolbook with v^3 = xi
        Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
        Fq2 r0 = t0 + ((c1 * o.c2) + (c2 * o.c1)).mul_by_xi();
        Fq2 r1 = (c0 * o.c1) + (c1 * o.c0) + t2.mul_by_xi();
        Fq2 r2 = (c0 * o.c2) + t1 + (c2 * o.c0);
        return {r0, r1, r2};
    }
    Fq6 mul_by_v() const { return {c2.mul_by_xi(), c0, c1}; }
    Fq6 inv() const {
        Fq2 a = c0.sqr() - (c1 * c2).mul_by_xi();
        Fq2 b = c2.sqr().mul_by_xi() - (c0 * c1);
        Fq2 c = c1.sqr() - (c0 * c2);
        Fq2 d = ((c2 * b) + (c1 * c)).mul_by_xi() + (c0 * a);
        Fq2 di = d.inv();
        return {a * di, b * di, c * di};
    }
};

struct Fq12 {
    Fq6 c0, c1;  // c0 + c1 w
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq12 operator*(const Fq12& o) const {  // w^2 = v
        Fq6 aa = c0 * o.c0, bb = c1 * o.c1;
        return {aa + bb.mul_by_v(), (c0 + c1) * (o.c0 + o.c1) - aa - bb};
    }
    Fq12 sqr() const { return (*this) * (*this); }
    Fq12 conj() const { return {c0, c1.neg()}; }  // = x^(q^6)
    Fq12 inv() const {
        Fq6 d = (c0 * c0 - (c1 * c1).mul_by_v()).inv();
        return {c0 * d, (c1 * d).neg()};
    }
    // x^(q^2): the coefficient of w^k (k = 0..5; Fq2 coefficients are fixed by q^2) picks up delta^k, delta = w^(q^2 - 1)
    Fq12 frobenius2(const Fq2 delta_pow[6]) const {
        return {{c0.c0, c0.c1 * delta_pow[2], c0.c2 * delta_pow[4]}, {c1.c0 * delta_pow[1], c1.c1 * delta_pow[3], c1.c2 * delta_pow[5]}};
    }
    Fq12 pow(const uint64_t* e, int limbs) const {
        Fq12 acc = one();
        for (int i = limbs - 1; i >= 0; --i)
            for (int b = 63; b >= 0; --b) {
                acc = acc.sqr();
                if ((e[i] >> b) & 1) acc = acc * (*this);
            }
        return acc;
    }
};

// layouts match halo2curves: raw Montgomery limbs, identity = all-zero coordinates
struct G1Point {
    Fq x, y;
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
};
struct G2Point {
    Fq2 x, y;
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
    G2Point neg() const { return {x, y.neg()}; }
};

namespace detail {
inline const uint64_t* exp_q_minus_1_over_6() {
    static const uint64_t e[4] = {0x34b017592414d4e1ull, 0xee9591c2e6bda1c2ull, 0xf40d60f3c0403964ull, 0x0810b7bdd032f006ull};
    return e;
}
inline const uint64_t* exp_q2_minus_1_over_6() {
    static const uint64_t e[8] = {0x348e0ec5b13a3c48ull, 0xc655abdcd6fc7580ull, 0x0c62aec4bcee7724ull, 0x2b66c518e9adb5ccull,
                                 0x5bd25464b3767342ull, 0x72ac96382e5e8e56ull, 0x0eef1294ab36cdafull, 0x01864b7413b4ca9aull};
    return e;
}
inline const uint64_t* exp_hard() {  // (q^4 - q^2 + 1) / r
    static const uint64_t e[12] = {0xe81bb482ccdf42b1ull, 0x5abf5cc4f49c36d4ull, 0xf1154e7e1da014fdull, 0xdcc7b44c87cdbacfull,
                                   0xaaa441e3954bcf8aull, 0x6b887d56d5095f23ull, 0x79581e16f3fd90c6ull, 0x3b1b1355d189227dull,
                                   0x4e529a5861876f6bull, 0x6c0eb522d5b12278ull, 0x331ec15183177fafull, 0x01baaa710b0759adull};
    return e;
}
inline Fq fq_small(uint32_t v) {
    Fq t = Fq::zero();
    t.l.v[0] = v;
    return t.to_mont();
}
inline Fq2 xi() { return {fq_small(9), Fq::one()}; }

struct Consts {
    Fq2 gamma2, gamma3;  // gamma = xi^((q-1)/6) = w^(q-1): the twist-coordinate Frobenius is (conj(x) gamma^2, conj(y) gamma^3)
    Fq2 delta_pow[6];    // delta = xi^((q^2-1)/6) = w^(q^2-1)
    Consts() {
        Fq2 g = xi().pow(exp_q_minus_1_over_6(), 4);
        gamma2 = g.sqr();
        gamma3 = gamma2 * g;
        Fq2 d = xi().pow(exp_q2_minus_1_over_6(), 8);
        delta_pow[0] = Fq2::one();
        for (int k = 1; k < 6; ++k) delta_pow[k] = delta_pow[k - 1] * d;
    }
};
inline const Consts& consts() {
    static const Consts c;
    return c;
}

// the line through T (slope lambda on the twist) evaluated at P, in the tower basis:
//   -yP  +  (lambda xP) w  +  (yT - lambda xT) w^3,    w^3 = v w
inline Fq12 line(const Fq2& lambda, const G2Point& t, const G1Point& p) {
    Fq12 l;
    l.c0 = {Fq2{p.y.neg(), Fq::zero()}, Fq2::zero(), Fq2::zero()};
    l.c1 = {lambda.scale(p.x), t.y - lambda * t.x, Fq2::zero()};
    return l;
}
// the vertical line x - xT at P: xP - xT w^2  (w^2 = v)
inline Fq12 vertical(const G2Point& t, const G1Point& p) {
    Fq12 l;
    l.c0 = {Fq2{p.x, Fq::zero()}, t.x.neg(), Fq2::zero()};
    l.c1 = Fq6::zero();
    return l;
}
// f *= line(T, Q)(P); T += Q   (handles T == Q, T == -Q, identities)
inline void step(Fq12& f, G2Point& t, bool& t_inf, const G2Point& q, const G1Point& p) {
    if (t_inf) {  // the line is the constant 1 up to a factor killed by the final exponentiation
        if (&q == &t) return;  // doubling the point at infinity: T stays at infinity (q aliases t in the doubling step)
        t = q;
        t_inf = false;
        return;
    }
    Fq2 lambda;
    if (t.x == q.x) {
        if (!(t.y == q.y) || t.y.is_zero()) {  // T = -Q: vertical line, T + Q = O
            f = f * vertical(t, p);
            t_inf = true;
            return;
        }
        Fq2 x2 = t.x.sqr();
        lambda = (x2.dbl() + x2) * t.y.dbl().inv();
    } else {
        lambda = (q.y - t.y) * (q.x - t.x).inv();
    }
    f = f * line(lambda, t, p);
    Fq2 x3 = lambda.sqr() - t.x - q.x;
    Fq2 y3 = lambda * (t.x - x3) - t.y;
    t = {x3, y3};
}
inline G2Point frobenius_twist(const G2Point& q) {
    const Consts& c = consts();
    return {q.x.conj() * c.gamma2, q.y.conj() * c.gamma3};
}
}  // namespace detail

// f_{6t+2, Q}(P) * the two Frobenius lines; 1 if either point is the identity
inline Fq12 miller_loop(const G1Point& p, const G2Point& q) {
    Fq12 f = Fq12::one();
    if (p.is_identity() || q.is_identity()) return f;
    const uint64_t ate_low = 0x9d797039be763ba8ull;  // 6t + 2 = 2^64 + ate_low, t = 4965661367192848881
    G2Point t = q;
    bool t_inf = false;
    for (int i = 63; i >= 0; --i) {
        f = f.sqr();
        detail::step(f, t, t_inf, t, p);
        if ((ate_low >> i) & 1) detail::step(f, t, t_inf, q, p);
    }
    G2Point q1 = detail::frobenius_twist(q);
    G2Point nq2 = detail::frobenius_twist(q1).neg();
    detail::step(f, t, t_inf, q1, p);
    detail::step(f, t, t_inf, nq2, p);
    return f;
}

inline Fq12 final_exponentiation(const Fq12& f) {
    Fq12 a = f.conj() * f.inv();                                     // f^(q^6 - 1)
    Fq12 b = a.frobenius2(detail::consts().delta_pow) * a;           // ^(q^2 + 1)
    return b.pow(detail::exp_hard(), 12);                            // ^((q^4 - q^2 + 1) / r)
}

inline Fq12 pairing(const G1Point& p, const G2Point& q) { return final_exponentiation(miller_loop(p, q)); }

// prod_i e(P_i, Q_i) == 1 for TRUSTED, already validated points (no on-curve / subgroup checks here: see
// pairing_check_validated below for the full EIP-197 input validation)
inline bool pairing_check(const std::vector<std::pair<G1Point, G2Point>>& pairs) {
    Fq12 f = Fq12::one();
    for (const auto& pr : pairs) f = f * miller_loop(pr.first, pr.second);
    return final_exponentiation(f) == Fq12::one();
}

// ---- encodings ------------------------------------------------------------------------------------------------
// canonical big-endian 32 B -> Montgomery Fq; false if >= q
inline bool fq_from_be32(const uint8_t* be, Fq* out) {
    Fq t;
    for (int i = 0; i < 8; ++i) {
        const uint8_t* p = be + 32 - 4 * (i + 1);
        t.l.v[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
    uint32_t m[8], d[8];
    Fq::modulus(m);
    if (!b200zk::leaf::sub8(d, t.l.v, m)) return false;  // sub8 reports the borrow: none means t >= q
    *out = t.to_mont();
    return true;
}
// EIP-197 G2 encoding (x_c1, x_c0, y_c1, y_c0), 128 B big-endian -- the word order of evm_verifier.yul:1230-1239
inline bool g2_from_eip197(const uint8_t* be128, G2Point* out) {
    return fq_from_be32(be128, &out->x.c1) && fq_from_be32(be128 + 32, &out->x.c0) && fq_from_be32(be128 + 64, &out->y.c1) &&
           fq_from_be32(be128 + 96, &out->y.c0);
}
// The KZG accumulator as the aggregation circuits expose it: 12 big-endian 32 B words holding limbs of 88 bits,
// (lhs.x, lhs.y, rhs.x, rhs.y) x 3 limbs, least significant first (/root/reference/integration/tests/unit_tests.rs:32,
// /root/reference/integration/configs/layer6.config:12-13 `limb_bits 88, num_limbs 3`).
inline bool accumulator_from_limbs(const uint8_t* be384, G1Point* lhs, G1Point* rhs) {
    Fq* dst[4] = {&lhs->x, &lhs->y, &rhs->x, &rhs->y};
    for (int c = 0; c < 4; ++c) {
        uint8_t be[32] = {0};  // value = l0 + l1 << 88 + l2 << 176, assembled big-endian
        for (int l = 0; l < 3; ++l) {
            const uint8_t* w = be384 + 32 * (3 * c + l);
            for (int i = 0; i < 21; ++i)
                if (w[i]) return false;  // a limb is < 2^88: the upper 21 bytes are zero
            // limb l occupies bits [88 l, 88 l + 88): bytes 31 - 11 l - 10 .. 31 - 11 l of the big-endian value
            for (int i = 0; i < 11; ++i) {
                int pos = 31 - 11 * l - i;
                if (pos < 0) {
                    if (w[31 - i]) return false;
                    continue;
                }
                be[pos] = w[31 - i];
            }
        }
        if (!fq_from_be32(be, dst[c])) return false;
    }
    return true;
}
inline bool g1_on_curve(const G1Point& p) {
    if (p.is_identity()) return true;
    return p.y * p.y == p.x * p.x * p.x + detail::fq_small(3);
}
inline bool g2_on_curve(const G2Point& q) {
    if (q.is_identity()) return true;
    Fq2 b2 = Fq2{detail::fq_small(3), Fq::zero()} * detail::xi().inv();
    return q.y.sqr() == q.x.sqr() * q.x + b2;
}

// ---- G2 group law on the twist (affine; host-side tooling: [tau]G2 of a test SRS, the subgroup check) ----------------
inline G2Point g2_add(const G2Point& a, const G2Point& b) {
    if (a.is_identity()) return b;
    if (b.is_identity()) return a;
    Fq2 lambda;
    if (a.x == b.x) {
        if (!(a.y == b.y) || a.y.is_zero()) return {Fq2::zero(), Fq2::zero()};
        Fq2 x2 = a.x.sqr();
        lambda = (x2.dbl() + x2) * a.y.dbl().inv();
    } else {
        lambda = (b.y - a.y) * (b.x - a.x).inv();
    }
    Fq2 x3 = lambda.sqr() - a.x - b.x;
    return {x3, lambda * (a.x - x3) - a.y};
}
// [s]Q, s given as 4 x u64 little-endian CANONICAL limbs (not Montgomery)
inline G2Point g2_mul(const G2Point& q, const uint64_t s[4]) {
    G2Point acc{Fq2::zero(), Fq2::zero()};
    for (int i = 3; i >= 0; --i)
        for (int b = 63; b >= 0; --b) {
            acc = g2_add(acc, acc);
            if ((s[i] >> b) & 1) acc = g2_add(acc, q);
        }
    return acc;
}
inline G2Point g2_generator() {  // halo2curves G2 generator = the constants of evm_verifier.yul:1230-1233
    static const uint8_t be[128] = {
        0x19, 0x8e, 0x93, 0x93, 0x92, 0x0d, 0x48, 0x3a, 0x72, 0x60, 0xbf, 0xb7, 0x31, 0xfb, 0x5d, 0x25, 0xf1, 0xaa, 0x49, 0x33, 0x35, 0xa9, 0xe7, 0x12, 0x97, 0xe4, 0x85, 0xb7, 0xae, 0xf3, 0x12, 0xc2,
        0x18, 0x00, 0xde, 0xef, 0x12, 0x1f, 0x1e, 0x76, 0x42, 0x6a, 0x00, 0x66, 0x5e, 0x5c, 0x44, 0x79, 0x67, 0x43, 0x22, 0xd4, 0xf7, 0x5e, 0xda, 0xdd, 0x46, 0xde, 0xbd, 0x5c, 0xd9, 0x92, 0xf6, 0xed,
        0x09, 0x06, 0x89, 0xd0, 0x58, 0x5f, 0xf0, 0x75, 0xec, 0x9e, 0x99, 0xad, 0x69, 0x0c, 0x33, 0x95, 0xbc, 0x4b, 0x31, 0x33, 0x70, 0xb3, 0x8e, 0xf3, 0x55, 0xac, 0xda, 0xdc, 0xd1, 0x22, 0x97, 0x5b,
        0x12, 0xc8, 0x5e, 0xa5, 0xdb, 0x8c, 0x6d, 0xeb, 0x4a, 0xab, 0x71, 0x80, 0x8d, 0xcb, 0x40, 0x8f, 0xe3, 0xd1, 0xe7, 0x69, 0x0c, 0x43, 0xd3, 0x7b, 0x4c, 0xe6, 0xcc, 0x01, 0x66, 0xfa, 0x7d, 0xaa};
    G2Point g;
    g2_from_eip197(be, &g);
    return g;
}
// Q in the r-torsion: [r]Q == O (G2's cofactor is not 1, so on-curve alone does not put a point in the pairing's domain)
inline bool g2_in_subgroup(const G2Point& q) {
    static const uint64_t r[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
    return g2_mul(q, r).is_identity();
}
// EIP-197 input validation in full: both points on their curves, the G2 point in the r-torsion
inline bool pairing_check_validated(const std::vector<std::pair<G1Point, G2Point>>& pairs, bool* valid_input = nullptr) {
    for (const auto& pr : pairs)
        if (!g1_on_curve(pr.first) || !g2_on_curve(pr.second) || !g2_in_subgroup(pr.second)) {
            if (valid_input) *valid_input = false;
            return false;
        }
    if (valid_input) *valid_input = true;
    return pairing_check(pairs);
}

// snark-verifier's KzgDecidingKey check for one accumulator: e(lhs, g2) * e(rhs, neg_s_g2) == 1
inline bool verify_kzg_accumulator(const G1Point& lhs, const G1Point& rhs, const G2Point& g2, const G2Point& neg_s_g2) {
    if (!g1_on_curve(lhs) || !g1_on_curve(rhs) || !g2_on_curve(g2) || !g2_on_curve(neg_s_g2)) return false;
    return pairing_check({{lhs, g2}, {rhs, neg_s_g2}});
}
// a single KZG opening: commitment C opens to y at x with witness W, against [tau]G2:
//   e(C - y G, G2) * e(-W, [tau]G2 - x G2) == 1, rearranged to avoid G2 arithmetic:  e(C - y G + x W, G2) * e(-W, [tau]G2) == 1
// (the caller supplies  lhs = C - y G + x W  and  W;  both are outputs of the device MSM / group ops)
inline bool verify_kzg_opening(const G1Point& c_minus_yg_plus_xw, const G1Point& w, const G2Point& g2, const G2Point& tau_g2) {
    if (!g1_on_curve(c_minus_yg_plus_xw) || !g1_on_curve(w) || !g2_on_curve(g2) || !g2_on_curve(tau_g2)) return false;
    G1Point nw = {w.x, w.y.neg()};
    return pairing_check({{c_minus_yg_plus_xw, g2}, {nw, tau_g2}});
}

}  // namespace pairing
}  // namespace halo2_b200
