// evm_verifier_b200.hpp — `EVMVerifier` (/root/reference/integration/src/verifier.rs:5-22): deploy the generated verifier contract and
// call it with the proof as calldata; the proof is valid iff the call does not revert.  The reference hands the deployment code
// (`evm_verifier.bin`, DEPLOYMENT_CODE_FILENAME) to an EVM (`prover::deploy_and_call`, revm); here the EVM is a small stack machine
// for exactly the opcodes that file contains, with the alt_bn128 precompiles served by the product's own host code:
//   0x05 modexp                     256-bit square-and-multiply (the program inverts field elements with it)
//   0x06 ecAdd, 0x07 ecMul          csrc/ec.cuh (XYZZ group law, the functions the device MSM uses, compiled for the host)
//   0x08 ecPairing                  pairing_bn254.hpp (EIP-197 validation: on-curve, G2 subgroup)
//   KECCAK256                       keccak256.hpp (the EVM transcript)
// Any other opcode raises (no silent mis-execution).  Gas is not metered (GAS pushes 2^256 - 1).  The calldata of the layer-6
// ("bundle") proof is proof.data with pi.data spliced in after the 12 accumulator limbs
// (/root/reference/integration/tests/unit_tests.rs:30-32) -- `calldata_of()`.
// Host-only, header-only.  tests/test_evm_verifier_kat.py runs release-v0.13.1/evm_verifier.bin on the shipped proof with this
// machine next to the Python one (tests/evm_bytecode.py): same verdicts, same precompile and hash call counts.
#pragma once
#include <cstring>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include "csrc/ec.cuh"
#include "keccak256.hpp"
#include "pairing_bn254.hpp"

namespace halo2_b200 {
namespace evm {

struct U256 {  // little-endian 64-bit limbs
    uint64_t w[4] = {0, 0, 0, 0};
    static U256 from_u64(uint64_t v) { U256 r; r.w[0] = v; return r; }
    static U256 from_be(const uint8_t* be, size_t len = 32) {  // up to 32 big-endian bytes
        U256 r;
        for (size_t i = 0; i < len; ++i) r.w[(len - 1 - i) / 8] |= (uint64_t)be[i] << (8 * ((len - 1 - i) % 8));
        return r;
    }
    void to_be(uint8_t* be) const {
        for (int i = 0; i < 32; ++i) be[i] = (uint8_t)(w[(31 - i) / 8] >> (8 * ((31 - i) % 8)));
    }
    bool is_zero() const { return (w[0] | w[1] | w[2] | w[3]) == 0; }
    bool fits_u64() const { return (w[1] | w[2] | w[3]) == 0; }
    bool operator==(const U256& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
    bool operator<(const U256& o) const {
        for (int i = 3; i >= 0; --i)
            if (w[i] != o.w[i]) return w[i] < o.w[i];
        return false;
    }
};
inline U256 add(const U256& a, const U256& b, bool* carry = nullptr) {
    U256 r;
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (unsigned __int128)a.w[i] + b.w[i];
        r.w[i] = (uint64_t)c;
        c >>= 64;
    }
    if (carry) *carry = c != 0;
    return r;
}
inline U256 sub(const U256& a, const U256& b) {  // mod 2^256
    U256 r;
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - borrow;
        r.w[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return r;
}
inline U256 shl(const U256& v, uint64_t s) {
    U256 r;
    if (s >= 256) return r;
    const int limbs = (int)(s / 64), bits = (int)(s % 64);
    for (int i = 3; i >= limbs; --i) {
        r.w[i] = v.w[i - limbs] << bits;
        if (bits && i - limbs - 1 >= 0) r.w[i] |= v.w[i - limbs - 1] >> (64 - bits);
    }
    return r;
}
// value of `n` limbs (little-endian) modulo m, m != 0: shift-and-subtract long division
inline U256 reduce(const uint64_t* limbs, int n, const U256& m) {
    U256 r;
    for (int bit = 64 * n - 1; bit >= 0; --bit) {
        const bool top = r.w[3] >> 63;
        r = shl(r, 1);
        r.w[0] |= (limbs[bit / 64] >> (bit % 64)) & 1;
        if (top || !(r < m)) r = sub(r, m);
    }
    return r;
}
inline U256 mod(const U256& a, const U256& m) { return m.is_zero() ? U256() : reduce(a.w, 4, m); }
inline U256 addmod(const U256& a, const U256& b, const U256& m) {
    if (m.is_zero()) return U256();
    bool carry;
    U256 s = add(a, b, &carry);
    uint64_t l[5] = {s.w[0], s.w[1], s.w[2], s.w[3], carry ? 1ull : 0ull};
    return reduce(l, 5, m);
}
inline U256 mulmod(const U256& a, const U256& b, const U256& m) {
    if (m.is_zero()) return U256();
    uint64_t p[8] = {0};
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (unsigned __int128)a.w[i] * b.w[j] + p[i + j];
            p[i + j] = (uint64_t)c;
            c >>= 64;
        }
        p[i + 4] = (uint64_t)c;
    }
    return reduce(p, 8, m);
}
inline U256 powmod(const U256& base, const U256& exp, const U256& m) {
    if (m.is_zero()) return U256();
    U256 acc = mod(U256::from_u64(1), m), b = mod(base, m);
    for (int bit = 255; bit >= 0; --bit) {
        acc = mulmod(acc, acc, m);
        if ((exp.w[bit / 64] >> (bit % 64)) & 1) acc = mulmod(acc, b, m);
    }
    return acc;
}

// ---- alt_bn128 precompiles on the product's host curve / pairing code; false = the precompile rejects its input
namespace precompile {
using pairing::G1Point;
inline bool g1_from_be64(const uint8_t* be, G1Point* p) {
    return pairing::fq_from_be32(be, &p->x) && pairing::fq_from_be32(be + 32, &p->y) && pairing::g1_on_curve(*p);
}
inline void g1_to_be64(const b200zk::XYZZ& v, uint8_t* out) {
    b200zk::Affine a = b200zk::xyzz_to_affine(v);
    b200zk::Fq xs[2] = {a.x.from_mont(), a.y.from_mont()};
    for (int c = 0; c < 2; ++c)
        for (int i = 0; i < 8; ++i) {
            const uint32_t w = xs[c].l.v[7 - i];
            for (int b = 0; b < 4; ++b) out[32 * c + 4 * i + b] = (uint8_t)(w >> (24 - 8 * b));
        }
}
inline bool ec_add(const uint8_t* in128, uint8_t* out64) {
    G1Point p, q;
    if (!g1_from_be64(in128, &p) || !g1_from_be64(in128 + 64, &q)) return false;
    b200zk::XYZZ acc = b200zk::xyzz_from_affine(b200zk::Affine{p.x, p.y}), t = b200zk::xyzz_from_affine(b200zk::Affine{q.x, q.y});
    b200zk::xyzz_add(acc, t);
    g1_to_be64(acc, out64);
    return true;
}
inline bool ec_mul(const uint8_t* in96, uint8_t* out64) {
    G1Point p;
    if (!g1_from_be64(in96, &p)) return false;
    b200zk::XYZZ acc = b200zk::XYZZ::identity();
    if (!p.is_identity())
        for (int i = 0; i < 256; ++i) {  // the scalar is any 256-bit integer, most significant bit first
            acc = b200zk::xyzz_dbl(acc);
            if ((in96[64 + i / 8] >> (7 - i % 8)) & 1) b200zk::xyzz_madd(acc, p.x, p.y);
        }
    g1_to_be64(acc, out64);
    return true;
}
inline bool ec_pairing(const uint8_t* in, size_t len, uint8_t* out32) {
    if (len % 192) return false;
    std::vector<std::pair<G1Point, pairing::G2Point>> pairs;
    for (size_t off = 0; off < len; off += 192) {
        G1Point a;
        pairing::G2Point b;
        if (!pairing::fq_from_be32(in + off, &a.x) || !pairing::fq_from_be32(in + off + 32, &a.y) || !pairing::g2_from_eip197(in + off + 64, &b)) return false;
        pairs.push_back({a, b});
    }
    bool valid = true;
    const bool one = pairing::pairing_check_validated(pairs, &valid);
    if (!valid) return false;
    std::memset(out32, 0, 32);
    out32[31] = one ? 1 : 0;
    return true;
}
}  // namespace precompile

struct Outcome {
    bool success = false;             // the code stopped with RETURN / STOP (false: REVERT, INVALID, a bad jump)
    std::vector<uint8_t> returndata;
    uint64_t steps = 0, keccak_calls = 0;
    uint64_t precompile_calls[9] = {0};  // by address (5 modexp, 6 ecAdd, 7 ecMul, 8 ecPairing)
};

class Machine {
  public:
    Machine(const std::vector<uint8_t>& code, const std::vector<uint8_t>& calldata) : code_(code), calldata_(calldata) {
        for (size_t i = 0; i < code_.size(); ++i) {
            const uint8_t op = code_[i];
            if (op == 0x5B) jumpdests_.insert(i);
            if (op >= 0x60 && op <= 0x7F) i += op - 0x5F;
        }
    }
    Outcome run() {
        Outcome out;
        size_t pc = 0;
        auto pop = [&]() {
            if (st_.empty()) throw std::runtime_error("evm: stack underflow");
            U256 v = st_.back();
            st_.pop_back();
            return v;
        };
        auto push = [&](const U256& v) {
            if (st_.size() >= 1024) throw std::runtime_error("evm: stack overflow");
            st_.push_back(v);
        };
        auto offset = [&](const U256& v) -> size_t {
            if (!v.fits_u64() || v.w[0] > (1ull << 32)) throw std::runtime_error("evm: memory offset out of range");
            return (size_t)v.w[0];
        };
        for (;;) {
            if (pc >= code_.size()) { out.success = true; return out; }
            const uint8_t op = code_[pc++];
            ++out.steps;
            if (op >= 0x60 && op <= 0x7F) {
                const size_t n = op - 0x5F;
                uint8_t buf[32] = {0};
                for (size_t i = 0; i < n && pc + i < code_.size(); ++i) buf[i] = code_[pc + i];
                push(U256::from_be(buf, n));
                pc += n;
            } else if (op >= 0x80 && op <= 0x8F) {
                const size_t n = op - 0x7F;
                if (st_.size() < n) throw std::runtime_error("evm: stack underflow");
                push(st_[st_.size() - n]);
            } else if (op >= 0x90 && op <= 0x9F) {
                const size_t n = op - 0x8F;
                if (st_.size() < n + 1) throw std::runtime_error("evm: stack underflow");
                std::swap(st_[st_.size() - 1], st_[st_.size() - 1 - n]);
            } else switch (op) {
                case 0x00: out.success = true; return out;
                case 0x01: { U256 a = pop(), b = pop(); push(add(a, b)); break; }
                case 0x03: { U256 a = pop(), b = pop(); push(sub(a, b)); break; }
                case 0x06: { U256 a = pop(), b = pop(); push(mod(a, b)); break; }
                case 0x08: { U256 a = pop(), b = pop(), m = pop(); push(addmod(a, b, m)); break; }
                case 0x09: { U256 a = pop(), b = pop(), m = pop(); push(mulmod(a, b, m)); break; }
                case 0x10: { U256 a = pop(), b = pop(); push(U256::from_u64(a < b)); break; }
                case 0x14: { U256 a = pop(), b = pop(); push(U256::from_u64(a == b)); break; }
                case 0x15: push(U256::from_u64(pop().is_zero())); break;
                case 0x16: { U256 a = pop(), b = pop(); for (int i = 0; i < 4; ++i) a.w[i] &= b.w[i]; push(a); break; }
                case 0x17: { U256 a = pop(), b = pop(); for (int i = 0; i < 4; ++i) a.w[i] |= b.w[i]; push(a); break; }
                case 0x1B: { U256 s = pop(), v = pop(); push(s.fits_u64() ? shl(v, s.w[0]) : U256()); break; }
                case 0x20: {
                    const size_t p = offset(pop()), n = offset(pop());
                    grow(p + n);
                    const auto d = hash::keccak256(mem_.data() + p, n);
                    ++out.keccak_calls;
                    push(U256::from_be(d.data()));
                    break;
                }
                case 0x35: {
                    const U256 pv = pop();
                    uint8_t buf[32] = {0};
                    if (pv.fits_u64())
                        for (size_t i = 0; i < 32; ++i)
                            if (pv.w[0] + i < calldata_.size()) buf[i] = calldata_[pv.w[0] + i];
                    push(U256::from_be(buf));
                    break;
                }
                case 0x39: {  // CODECOPY(dest, offset, size)
                    const size_t d = offset(pop()), o = offset(pop()), n = offset(pop());
                    grow(d + n);
                    for (size_t i = 0; i < n; ++i) mem_[d + i] = o + i < code_.size() ? code_[o + i] : 0;
                    break;
                }
                case 0x50: pop(); break;
                case 0x51: { const size_t p = offset(pop()); grow(p + 32); push(U256::from_be(mem_.data() + p)); break; }
                case 0x52: { const size_t p = offset(pop()); const U256 v = pop(); grow(p + 32); v.to_be(mem_.data() + p); break; }
                case 0x53: { const size_t p = offset(pop()); const U256 v = pop(); grow(p + 1); mem_[p] = (uint8_t)v.w[0]; break; }
                case 0x56: {
                    const U256 d = pop();
                    if (!d.fits_u64() || !jumpdests_.count((size_t)d.w[0])) return out;
                    pc = (size_t)d.w[0];
                    break;
                }
                case 0x57: {
                    const U256 d = pop(), cond = pop();
                    if (!cond.is_zero()) {
                        if (!d.fits_u64() || !jumpdests_.count((size_t)d.w[0])) return out;
                        pc = (size_t)d.w[0];
                    }
                    break;
                }
                case 0x5A: { U256 g; g.w[0] = g.w[1] = g.w[2] = g.w[3] = ~0ull; push(g); break; }
                case 0x5B: break;
                case 0xF3: {
                    const size_t p = offset(pop()), n = offset(pop());
                    grow(p + n);
                    out.returndata.assign(mem_.begin() + p, mem_.begin() + p + n);
                    out.success = true;
                    return out;
                }
                case 0xFA: {  // STATICCALL(gas, addr, in, insize, out, outsize)
                    pop();
                    const U256 addr = pop();
                    const size_t ip = offset(pop()), in = offset(pop()), op_ = offset(pop()), on = offset(pop());
                    grow(ip + in);
                    std::vector<uint8_t> input(mem_.begin() + ip, mem_.begin() + ip + in), result;
                    bool ok = false;
                    if (addr.fits_u64() && addr.w[0] >= 5 && addr.w[0] <= 8) {
                        ++out.precompile_calls[addr.w[0]];
                        ok = precompiled(addr.w[0], input, &result);
                    }
                    if (ok) {
                        grow(op_ + on);
                        for (size_t i = 0; i < on && i < result.size(); ++i) mem_[op_ + i] = result[i];
                    }
                    push(U256::from_u64(ok));
                    break;
                }
                case 0xFD: case 0xFE: return out;
                default: throw std::runtime_error("evm: opcode 0x" + hex2(op) + " is not one the verifier uses");
            }
        }
    }

  private:
    static std::string hex2(uint8_t v) {
        const char* d = "0123456789abcdef";
        return std::string(1, d[v >> 4]) + d[v & 15];
    }
    void grow(size_t end) {
        if (end > mem_.size()) mem_.resize(end, 0);
    }
    static bool precompiled(uint64_t addr, std::vector<uint8_t> in, std::vector<uint8_t>* out) {
        if (addr == 5) {  // <len_b, len_e, len_m, b, e, m>, every length <= 32 here
            in.resize(std::max<size_t>(in.size(), 96), 0);
            const U256 lb = U256::from_be(in.data()), le = U256::from_be(in.data() + 32), lm = U256::from_be(in.data() + 64);
            if (!lb.fits_u64() || !le.fits_u64() || !lm.fits_u64() || lb.w[0] > 32 || le.w[0] > 32 || lm.w[0] > 32) return false;
            in.resize(96 + lb.w[0] + le.w[0] + lm.w[0], 0);
            const U256 b = U256::from_be(in.data() + 96, lb.w[0]), e = U256::from_be(in.data() + 96 + lb.w[0], le.w[0]),
                       m = U256::from_be(in.data() + 96 + lb.w[0] + le.w[0], lm.w[0]);
            uint8_t be[32];
            powmod(b, e, m).to_be(be);
            out->assign(be + 32 - lm.w[0], be + 32);
            return true;
        }
        if (addr == 6) {
            in.resize(128, 0);
            out->resize(64);
            return precompile::ec_add(in.data(), out->data());
        }
        if (addr == 7) {
            in.resize(96, 0);
            out->resize(64);
            return precompile::ec_mul(in.data(), out->data());
        }
        out->resize(32);
        return precompile::ec_pairing(in.data(), in.size(), out->data());
    }
    const std::vector<uint8_t>& code_;
    const std::vector<uint8_t>& calldata_;
    std::set<size_t> jumpdests_;
    std::vector<U256> st_;
    std::vector<uint8_t> mem_;
};

// `prover::deploy_and_call`: run the creation code, then call the runtime it returns with the calldata
inline std::vector<uint8_t> deploy(const std::vector<uint8_t>& creation_code) {
    const std::vector<uint8_t> none;
    Outcome o = Machine(creation_code, none).run();
    if (!o.success || o.returndata.empty()) throw std::runtime_error("evm: the creation code did not return a runtime");
    return o.returndata;
}
inline Outcome call(const std::vector<uint8_t>& runtime, const std::vector<uint8_t>& calldata) { return Machine(runtime, calldata).run(); }

// the calldata of an EVM proof: the accumulator limbs (12 words), the public-input words, the rest of the proof
inline std::vector<uint8_t> calldata_of(const std::vector<uint8_t>& proof, const std::vector<uint8_t>& pi) {
    if (proof.size() < 384) throw std::runtime_error("evm: an EVM proof starts with 12 accumulator words");
    std::vector<uint8_t> cd(proof.begin(), proof.begin() + 384);
    cd.insert(cd.end(), pi.begin(), pi.end());
    cd.insert(cd.end(), proof.begin() + 384, proof.end());
    return cd;
}

class EVMVerifier {  // verifier.rs: EVMVerifier(deployment_code).verify_evm_proof(call_data)
  public:
    explicit EVMVerifier(std::vector<uint8_t> deployment_code) : code_(std::move(deployment_code)) {}
    bool verify_evm_proof(const std::vector<uint8_t>& call_data, Outcome* detail = nullptr) const {
        try {
            const std::vector<uint8_t> runtime = deploy(code_);
            Outcome o = call(runtime, call_data);
            if (detail) *detail = o;
            return o.success;
        } catch (const std::exception&) {
            return false;
        }
    }

  private:
    std::vector<uint8_t> code_;
};

}  // namespace evm
}  // namespace halo2_b200
