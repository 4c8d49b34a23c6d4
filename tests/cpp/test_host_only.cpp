// CPU-only checks of the C++ host mirror (scroll-prover_b200/halo2_b200.hpp): everything here runs without a CUDA device --
// domain constants, the plonk program generators and their lowering through b200zk_graph_check.  Driven by
// tests/test_cpp_mirror.py, which passes the expected values (from the reference's chunk.protocol and from the Python twins).
//   argv: omega25 omega_inv25 n_inv25 (64 hex digits each, little-endian limb bytes)  perm_instr perm_slots  lookup_instr lookup_slots
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../scroll-prover_b200/halo2_b200.hpp"

using namespace halo2_b200;

#define REQUIRE(c)                                                     \
    do {                                                               \
        if (!(c)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

static Fr from_hex(const char* h) {
    Fr r;
    uint8_t b[32];
    for (int i = 0; i < 32; ++i) {
        unsigned v;
        std::sscanf(h + 2 * i, "%2x", &v);
        b[i] = (uint8_t)v;
    }
    std::memcpy(r.l, b, 32);
    return r;
}

int main(int argc, char** argv) {
    if (argc != 8) {
        std::printf("usage\n");
        return 2;
    }
    // EvaluationDomain::new(5, 25) as the reference dumped it (release-v0.13.1/chunk.protocol "domain")
    EvaluationDomain dom = EvaluationDomain::new_(5, 25);
    REQUIRE(dom.k == 25 && dom.extended_k == 27 && dom.quotient_poly_degree == 4);
    REQUIRE(dom.omega == from_hex(argv[1]));
    REQUIRE(dom.omega_inv == from_hex(argv[2]));
    REQUIRE(dom.ifft_divisor == from_hex(argv[3]));

    using plonk::ValueSource;
    // evaluate_h's permutation section: 3 sets over 8 columns in chunks of 3, last rotation -6
    {
        plonk::GraphEvaluator ev;
        uint32_t r0 = ev.add_rotation(0);
        std::vector<ValueSource> z, v, s;
        for (uint32_t i = 0; i < 3; ++i) z.push_back(ValueSource::Advice(i, r0));
        for (uint32_t j = 0; j < 8; ++j) v.push_back(ValueSource::Advice(3 + j, r0));
        for (uint32_t j = 0; j < 8; ++j) s.push_back(ValueSource::Fixed(j, r0));
        Fr delta = detail::from_dev(detail::from_u64(7).pow_u64(1ull << 28));
        plonk::permutation_constraints(ev, z, 3, v, s, ValueSource::Fixed(8, r0), ValueSource::Fixed(9, r0), ValueSource::Fixed(10, r0), -6, delta);
        auto [ni, ns] = ev.check();
        std::printf("perm %u %u\n", ni, ns);
        REQUIRE(ni == (uint32_t)std::atoi(argv[4]) && ns == (uint32_t)std::atoi(argv[5]));
    }
    // one log-derivative lookup with 3 inputs
    {
        plonk::GraphEvaluator ev;
        uint32_t r0 = ev.add_rotation(0);
        plonk::lookup_constraints(ev, {ValueSource::Advice(0, r0), ValueSource::Advice(1, r0), ValueSource::Advice(2, r0)}, ValueSource::Advice(3, r0),
                                  ValueSource::Advice(4, r0), ValueSource::Advice(5, r0), ValueSource::Fixed(0, r0), ValueSource::Fixed(1, r0),
                                  ValueSource::Fixed(2, r0));
        auto [ni, ns] = ev.check();
        std::printf("lookup %u %u\n", ni, ns);
        REQUIRE(ni == (uint32_t)std::atoi(argv[6]) && ns == (uint32_t)std::atoi(argv[7]));
    }
    // a malformed program is reported with its reason, not a crash
    {
        plonk::GraphEvaluator ev;
        ev.add(B200ZK_CALC_ADD, ValueSource::Intermediate(5), ValueSource::Constant(0));
        bool panicked = false;
        try {
            ev.check();
        } catch (const Panic& e) {
            panicked = std::strstr(e.what(), "earlier calculation") != nullptr;
        }
        REQUIRE(panicked);
    }
    std::printf("HOST OK\n");
    return 0;
}
