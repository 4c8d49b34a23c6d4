// Driver for tests/test_evm_verifier_kat.py: scroll-prover_b200/evm_verifier_b200.hpp (the product's EVMVerifier mirror) on files
//   test_evm_verifier <deployment code> <proof.data> <pi.data> [byte position to flip in the calldata | -N to drop the last N bytes]
//   test_evm_verifier --arith <a> <b> <m>   (64 hex digits each): addmod, mulmod, a mod m, a^b mod m, a + b, a - b, a << (b mod 2^64) -- one per line
// prints `ACCEPT|REJECT runtime_bytes=.. steps=.. keccak=.. modexp=.. ecadd=.. ecmul=.. pairing=..`.  Host only: no device, no libb200zk.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>

#include "../../scroll-prover_b200/evm_verifier_b200.hpp"

using namespace halo2_b200;

static std::vector<uint8_t> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

static evm::U256 parse(const char* h) {
    uint8_t be[32] = {0};
    for (int i = 0; i < 32; ++i) {
        unsigned v = 0;
        std::sscanf(h + 2 * i, "%2x", &v);
        be[i] = (uint8_t)v;
    }
    return evm::U256::from_be(be);
}
static void show(const evm::U256& v) {
    uint8_t be[32];
    v.to_be(be);
    for (uint8_t b : be) std::printf("%02x", b);
    std::printf("\n");
}

int main(int argc, char** argv) {
    if (argc == 5 && std::string(argv[1]) == "--arith") {
        const evm::U256 a = parse(argv[2]), b = parse(argv[3]), m = parse(argv[4]);
        show(evm::addmod(a, b, m));
        show(evm::mulmod(a, b, m));
        show(evm::mod(a, m));
        show(evm::powmod(a, b, m));
        show(evm::add(a, b));
        show(evm::sub(a, b));
        show(b.fits_u64() ? evm::shl(a, b.w[0]) : evm::U256());
        return 0;
    }
    if (argc < 4) return 2;
    try {
        const std::vector<uint8_t> code = slurp(argv[1]);
        std::vector<uint8_t> calldata = evm::calldata_of(slurp(argv[2]), slurp(argv[3]));
        if (argc > 4) {
            const long pos = std::atol(argv[4]);
            if (pos < 0) calldata.resize(calldata.size() + pos);
            else calldata.at((size_t)pos) ^= 1;
        }
        const size_t runtime_bytes = evm::deploy(code).size();
        evm::Outcome o;
        const bool ok = evm::EVMVerifier(code).verify_evm_proof(calldata, &o);
        std::printf("%s runtime_bytes=%zu steps=%llu keccak=%llu modexp=%llu ecadd=%llu ecmul=%llu pairing=%llu\n", ok ? "ACCEPT" : "REJECT", runtime_bytes,
                    (unsigned long long)o.steps, (unsigned long long)o.keccak_calls, (unsigned long long)o.precompile_calls[5],
                    (unsigned long long)o.precompile_calls[6], (unsigned long long)o.precompile_calls[7], (unsigned long long)o.precompile_calls[8]);
        return 0;
    } catch (const std::exception& e) {
        std::printf("ERROR %s\n", e.what());
        return 1;
    }
}
