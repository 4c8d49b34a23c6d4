// C shim over scroll-prover_b200/serde_bn254.hpp for tests/test_serde_host.py
#include "../../scroll-prover_b200/serde_bn254.hpp"
using namespace halo2_b200::serde;

// decompress -> (x, y) Montgomery limbs (16 x u32); returns 1 on success
extern "C" int serde_host_decompress(const uint8_t* in32, uint32_t* out16) {
    G1Point p;
    if (!g1_from_compressed(in32, &p)) return 0;
    std::memcpy(out16, &p, 64);
    return 1;
}
extern "C" void serde_host_compress(const uint32_t* in16, uint8_t* out32) {
    G1Point p;
    std::memcpy(&p, in16, 64);
    g1_to_compressed(p, out32);
}
// parses a vk file; writes k, counts and the points (Montgomery limbs, fixed then permutation); returns 1 on success and
// re-serialises it into `rewritten` (same length) for a byte-exact round-trip check
extern "C" int serde_host_read_vk(const uint8_t* bytes, uint64_t len, uint32_t* k, uint32_t* n_fixed, uint32_t* n_perm, uint32_t* points,
                                  uint8_t* rewritten) {
    VerifyingKeyFile vk;
    if (!read_vk_processed(bytes, (size_t)len, &vk)) return 0;
    *k = vk.k;
    *n_fixed = (uint32_t)vk.fixed_commitments.size();
    *n_perm = (uint32_t)vk.permutation_commitments.size();
    size_t i = 0;
    for (const auto* v : {&vk.fixed_commitments, &vk.permutation_commitments})
        for (const auto& p : *v) std::memcpy(points + 16 * (i++), &p, 64);
    std::vector<uint8_t> w = write_vk_processed(vk);
    if (w.size() != len) return 0;
    std::memcpy(rewritten, w.data(), w.size());
    return 1;
}
