// keccak256.hpp — Keccak-256 with the original 0x01 .. 0x80 padding (the hash of the EVM and of scroll's public-input hashes).
// Used by proof_files.hpp (ChunkInfo::public_input_hash) and evm_verifier_b200.hpp (the KECCAK256 opcode: the EVM verifier's
// transcript).  Host-only, header-only; known answers in tests/test_snark_verifier_host.py.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace halo2_b200 {
namespace hash {

inline std::array<uint8_t, 32> keccak256(const uint8_t* data, size_t len) {
    static const uint64_t RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
                                    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
                                    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
                                    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
                                    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
    static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};  // [x + 5 y]
    uint64_t a[25] = {0};
    auto rol = [](uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; };
    auto permute = [&]() {
        for (int round = 0; round < 24; ++round) {
            uint64_t c[5], b[25];
            for (int x = 0; x < 5; ++x) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
            for (int x = 0; x < 5; ++x) {
                uint64_t d = c[(x + 4) % 5] ^ rol(c[(x + 1) % 5], 1);
                for (int y = 0; y < 5; ++y) a[x + 5 * y] ^= d;
            }
            for (int x = 0; x < 5; ++x)
                for (int y = 0; y < 5; ++y) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol(a[x + 5 * y], ROT[x + 5 * y]);
            for (int x = 0; x < 5; ++x)
                for (int y = 0; y < 5; ++y) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
            a[0] ^= RC[round];
        }
    };
    const size_t rate = 136;
    std::vector<uint8_t> msg(data, data + len);
    msg.push_back(0x01);
    while (msg.size() % rate) msg.push_back(0);
    msg.back() |= 0x80;
    for (size_t off = 0; off < msg.size(); off += rate) {
        for (size_t i = 0; i < rate / 8; ++i) {
            uint64_t w = 0;
            for (int b = 7; b >= 0; --b) w = (w << 8) | msg[off + 8 * i + b];
            a[i] ^= w;
        }
        permute();
    }
    std::array<uint8_t, 32> out;
    for (int i = 0; i < 4; ++i)
        for (int b = 0; b < 8; ++b) out[8 * i + b] = (uint8_t)(a[i] >> (8 * b));
    return out;
}

}  // namespace hash
}  // namespace halo2_b200
