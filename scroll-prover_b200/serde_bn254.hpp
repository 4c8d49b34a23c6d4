// serde_bn254.hpp — host-side point codecs and the verifying-key file layout.
//
// halo2curves' compressed G1 encoding (`SerdeFormat::Processed`): 32 bytes, little-endian canonical x; bit 254 carries the
// least significant bit of canonical y; bit 255 flags the identity.  `vk_*.vkey` files
// (/root/reference/release-v0.13.1/vk_chunk.vkey and friends, written by snark-verifier-sdk in Processed format):
// k as u32 BIG-endian, the number of fixed commitments as u32 BIG-endian, then that many compressed points, then the
// permutation commitments (compressed) up to the end of the file.  Host code on the host build of csrc/ff.cuh; validated
// in tests/test_serde_host.py against the reference's shipped verifying keys and the points of chunk.protocol.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "csrc/ff.cuh"

namespace halo2_b200 {
namespace serde {

using Fq = b200zk::Fq;
struct G1Point {
    Fq x, y;  // Montgomery limbs; identity = (0, 0)
    bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

inline Fq fq_small(uint32_t v) {
    Fq t = Fq::zero();
    t.l.v[0] = v;
    return t.to_mont();
}
// canonical little-endian bytes <-> Montgomery element
inline bool fq_from_le32(const uint8_t* le, Fq* out) {
    Fq t;
    for (int i = 0; i < 8; ++i) t.l.v[i] = (uint32_t)le[4 * i] | ((uint32_t)le[4 * i + 1] << 8) | ((uint32_t)le[4 * i + 2] << 16) | ((uint32_t)le[4 * i + 3] << 24);
    uint32_t m[8], d[8];
    Fq::modulus(m);
    if (!b200zk::leaf::sub8(d, t.l.v, m)) return false;  // no borrow: not reduced
    *out = t.to_mont();
    return true;
}
inline void fq_to_le32(const Fq& a, uint8_t* le) {
    Fq c = a.from_mont();
    for (int i = 0; i < 8; ++i)
        for (int b = 0; b < 4; ++b) le[4 * i + b] = (uint8_t)(c.l.v[i] >> (8 * b));
}
// square root for q = 3 (mod 4): a^((q + 1) / 4); false when a is not a square
inline bool fq_sqrt(const Fq& a, Fq* out) {
    uint32_t e[8];
    Fq::modulus(e);
    // (q + 1) / 4: q + 1 does not overflow 256 bits (q < 2^254)
    uint64_t carry = 1;
    for (int i = 0; i < 8; ++i) {
        uint64_t s = (uint64_t)e[i] + carry;
        e[i] = (uint32_t)s;
        carry = s >> 32;
    }
    for (int i = 0; i < 8; ++i) e[i] = (e[i] >> 2) | (i < 7 ? (e[i + 1] << 30) : 0);
    Fq acc = Fq::one();
    for (int i = 7; i >= 0; --i)
        for (int b = 31; b >= 0; --b) {
            acc = acc.sqr();
            if ((e[i] >> b) & 1) acc = acc * a;
        }
    if (!(acc.sqr() == a)) return false;
    *out = acc;
    return true;
}

inline void g1_to_compressed(const G1Point& p, uint8_t out[32]) {
    if (p.is_identity()) {
        std::memset(out, 0, 32);
        out[31] = 0x80;
        return;
    }
    uint8_t yb[32];
    fq_to_le32(p.x, out);
    fq_to_le32(p.y, yb);
    out[31] |= (uint8_t)((yb[0] & 1) << 6);
}
inline bool g1_from_compressed(const uint8_t in[32], G1Point* out) {
    uint8_t t[32];
    std::memcpy(t, in, 32);
    const bool inf = t[31] & 0x80, sign = t[31] & 0x40;
    t[31] &= 0x3f;
    if (inf) {
        for (int i = 0; i < 32; ++i)
            if (t[i]) return false;
        if (sign) return false;
        out->x = Fq::zero();
        out->y = Fq::zero();
        return true;
    }
    Fq x, y;
    if (!fq_from_le32(t, &x)) return false;
    if (!fq_sqrt(x.sqr() * x + fq_small(3), &y)) return false;  // y^2 = x^3 + 3
    uint8_t yb[32];
    fq_to_le32(y, yb);
    if (((yb[0] & 1) != 0) != sign) y = y.neg();
    out->x = x;
    out->y = y;
    return true;
}

struct VerifyingKeyFile {
    uint32_t k = 0;
    std::vector<G1Point> fixed_commitments, permutation_commitments;
};
inline bool read_vk_processed(const uint8_t* bytes, size_t len, VerifyingKeyFile* out) {
    auto be32 = [](const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; };
    if (len < 8 || (len - 8) % 32) return false;
    out->k = be32(bytes);
    const uint32_t n_fixed = be32(bytes + 4);
    const size_t n_points = (len - 8) / 32;
    if (n_fixed > n_points) return false;
    out->fixed_commitments.clear();
    out->permutation_commitments.clear();
    for (size_t i = 0; i < n_points; ++i) {
        G1Point p;
        if (!g1_from_compressed(bytes + 8 + 32 * i, &p)) return false;
        (i < n_fixed ? out->fixed_commitments : out->permutation_commitments).push_back(p);
    }
    return true;
}
inline std::vector<uint8_t> write_vk_processed(const VerifyingKeyFile& vk) {
    std::vector<uint8_t> out(8 + 32 * (vk.fixed_commitments.size() + vk.permutation_commitments.size()));
    auto put = [&](size_t off, uint32_t v) { out[off] = v >> 24; out[off + 1] = v >> 16; out[off + 2] = v >> 8; out[off + 3] = v; };
    put(0, vk.k);
    put(4, (uint32_t)vk.fixed_commitments.size());
    size_t off = 8;
    for (const auto* v : {&vk.fixed_commitments, &vk.permutation_commitments})
        for (const auto& p : *v) {
            g1_to_compressed(p, out.data() + off);
            off += 32;
        }
    return out;
}

}  // namespace serde
}  // namespace halo2_b200
