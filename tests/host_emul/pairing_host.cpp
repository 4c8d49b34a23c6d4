// C shim over scroll-prover_b200/pairing_bn254.hpp for the CPU tests (tests/test_pairing_host.py): the EIP-197 calling
// convention of precompile 0x08 (what release-v0.13.1/evm_verifier.yul:1240 calls), plus the accumulator decoder.
#include "../../scroll-prover_b200/pairing_bn254.hpp"

using namespace halo2_b200::pairing;

// input: n x 192 B = G1 (x, y big-endian) | G2 (x_c1, x_c0, y_c1, y_c0 big-endian).  returns 1 / 0, or -1 for a malformed point
extern "C" int pairing_host_eip197(const uint8_t* in, int n) {
    std::vector<std::pair<G1Point, G2Point>> pairs;
    for (int i = 0; i < n; ++i) {
        const uint8_t* p = in + 192 * (size_t)i;
        G1Point a;
        G2Point b;
        if (!fq_from_be32(p, &a.x) || !fq_from_be32(p + 32, &a.y) || !g2_from_eip197(p + 64, &b)) return -1;
        if (!g1_on_curve(a) || !g2_on_curve(b) || !g2_in_subgroup(b)) return -1;  // EIP-197 input validation
        pairs.push_back({a, b});
    }
    return pairing_check(pairs) ? 1 : 0;
}
// the accumulator words of proof.data / the instance column, decided against (g2, neg_s_g2) in EIP-197 encoding
extern "C" int pairing_host_accumulator(const uint8_t* be384, const uint8_t* g2_be128, const uint8_t* neg_s_g2_be128) {
    G1Point lhs, rhs;
    G2Point g2, x;
    if (!accumulator_from_limbs(be384, &lhs, &rhs) || !g2_from_eip197(g2_be128, &g2) || !g2_from_eip197(neg_s_g2_be128, &x)) return -1;
    return verify_kzg_accumulator(lhs, rhs, g2, x) ? 1 : 0;
}
// e(P, Q)^k == e(kP, Q) style checks need GT equality: compares the pairings of two (G1, G2) pairs
extern "C" int pairing_host_equal(const uint8_t* pair_a192, const uint8_t* pair_b192) {
    G1Point a1, b1;
    G2Point a2, b2;
    if (!fq_from_be32(pair_a192, &a1.x) || !fq_from_be32(pair_a192 + 32, &a1.y) || !g2_from_eip197(pair_a192 + 64, &a2)) return -1;
    if (!fq_from_be32(pair_b192, &b1.x) || !fq_from_be32(pair_b192 + 32, &b1.y) || !g2_from_eip197(pair_b192 + 64, &b2)) return -1;
    return pairing(a1, a2) == pairing(b1, b2) ? 1 : 0;
}

// ---- the EVM's alt_bn128 precompiles 0x06 (ecAdd) and 0x07 (ecMul) on the product's host curve code (csrc/ec.cuh compiled for
// the CPU, the very functions the device MSM uses): inputs / outputs big-endian coordinates, (0, 0) = identity; -1 for a point
// that is not on the curve.  Used by tests/test_evm_verifier_kat.py to run the reference's verifier program.
#include "../../scroll-prover_b200/csrc/ec.cuh"
static bool g1_from_be64(const uint8_t* be, G1Point* p) { return fq_from_be32(be, &p->x) && fq_from_be32(be + 32, &p->y) && g1_on_curve(*p); }
static void g1_to_be64(const b200zk::XYZZ& v, uint8_t* out) {
    b200zk::Affine a = b200zk::xyzz_to_affine(v);
    Fq xs[2] = {a.x.from_mont(), a.y.from_mont()};
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 8; ++i) {
            uint32_t w = xs[c].l.v[7 - i];
            out[32 * c + 4 * i + 0] = (uint8_t)(w >> 24);
            out[32 * c + 4 * i + 1] = (uint8_t)(w >> 16);
            out[32 * c + 4 * i + 2] = (uint8_t)(w >> 8);
            out[32 * c + 4 * i + 3] = (uint8_t)w;
        }
}
extern "C" int ec_host_add(const uint8_t* in128, uint8_t* out64) {
    G1Point p, q;
    if (!g1_from_be64(in128, &p) || !g1_from_be64(in128 + 64, &q)) return -1;
    b200zk::Affine a{p.x, p.y}, b{q.x, q.y};
    b200zk::XYZZ acc = b200zk::xyzz_from_affine(a), t = b200zk::xyzz_from_affine(b);
    b200zk::xyzz_add(acc, t);
    g1_to_be64(acc, out64);
    return 0;
}
extern "C" int ec_host_mul(const uint8_t* in96, uint8_t* out64) {
    G1Point p;
    if (!g1_from_be64(in96, &p)) return -1;
    b200zk::XYZZ acc = b200zk::XYZZ::identity();
    if (!p.is_identity())
        for (int i = 0; i < 256; ++i) {  // the scalar is any 256-bit integer (not reduced), most significant bit first
            acc = b200zk::xyzz_dbl(acc);
            if ((in96[64 + i / 8] >> (7 - i % 8)) & 1) b200zk::xyzz_madd(acc, p.x, p.y);
        }
    g1_to_be64(acc, out64);
    return 0;
}
