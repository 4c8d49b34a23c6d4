// Driver for tests/test_protocol_json.py: parses a `*.protocol` file with scroll-prover_b200/protocol_json.hpp, cross-checks the
// domain against the C++ mirror's EvaluationDomain::new_ and the preprocessed points against a `vk_*.vkey` file decoded by
// serde_bn254.hpp, and prints a JSON summary that the Python side compares with its own reading of the same file.
//   usage: test_protocol_json <file.protocol> [vk_file.vkey]
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../../scroll-prover_b200/halo2_b200.hpp"
#include "../../scroll-prover_b200/protocol_json.hpp"
#include "../../scroll-prover_b200/serde_bn254.hpp"

using namespace halo2_b200;

static std::string slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
static void print_limbs(const char* name, const protocol::Limbs4& v, bool comma = true) {
    std::printf("\"%s\": [%llu, %llu, %llu, %llu]%s", name, (unsigned long long)v.l[0], (unsigned long long)v.l[1], (unsigned long long)v.l[2],
                (unsigned long long)v.l[3], comma ? ", " : "");
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    try {
        protocol::PlonkProtocol p = protocol::parse_protocol(slurp(argv[1]));
        // domain == EvaluationDomain::new(j, k) of the mirror (j only affects the extended domain)
        EvaluationDomain dom = EvaluationDomain::new_(5, p.domain.k);
        bool dom_ok = std::memcmp(dom.omega.l, p.domain.gen.l, 32) == 0 && std::memcmp(dom.omega_inv.l, p.domain.gen_inv.l, 32) == 0 &&
                      std::memcmp(dom.ifft_divisor.l, p.domain.n_inv.l, 32) == 0 && dom.n == p.domain.n;
        int vk_match = -1;
        if (argc > 2) {
            std::string vk = slurp(argv[2]);
            serde::VerifyingKeyFile f;
            if (!serde::read_vk_processed((const uint8_t*)vk.data(), vk.size(), &f)) {
                std::printf("{\"error\": \"vk file does not parse\"}\n");
                return 1;
            }
            std::vector<serde::G1Point> pts = f.fixed_commitments;  // the protocol lists fixed then permutation commitments
            pts.insert(pts.end(), f.permutation_commitments.begin(), f.permutation_commitments.end());
            vk_match = (pts.size() == p.preprocessed.size() && f.k == p.domain.k) ? 1 : 0;
            for (size_t i = 0; vk_match == 1 && i < pts.size(); ++i)
                if (std::memcmp(pts[i].x.l.v, p.preprocessed[i].x.l, 32) != 0 || std::memcmp(pts[i].y.l.v, p.preprocessed[i].y.l, 32) != 0) vk_match = 0;
        }
        std::printf("{\"k\": %u, \"n\": %llu, ", p.domain.k, (unsigned long long)p.domain.n);
        print_limbs("n_inv", p.domain.n_inv);
        print_limbs("gen", p.domain.gen);
        std::printf("\"domain_matches_mirror\": %s, \"n_preprocessed\": %zu, ", dom_ok ? "true" : "false", p.preprocessed.size());
        if (!p.preprocessed.empty()) print_limbs("last_preprocessed_y", p.preprocessed.back().y);
        std::printf("\"num_instance\": [");
        for (size_t i = 0; i < p.num_instance.size(); ++i) std::printf("%s%llu", i ? ", " : "", (unsigned long long)p.num_instance[i]);
        std::printf("], \"num_witness\": [");
        for (size_t i = 0; i < p.num_witness.size(); ++i) std::printf("%s%llu", i ? ", " : "", (unsigned long long)p.num_witness[i]);
        std::printf("], \"num_challenge\": [");
        for (size_t i = 0; i < p.num_challenge.size(); ++i) std::printf("%s%llu", i ? ", " : "", (unsigned long long)p.num_challenge[i]);
        std::printf("], \"n_evaluations\": %zu, \"n_queries\": %zu, \"quotient\": [%llu, %llu], \"numerator_root\": \"%s\", ", p.evaluations.size(),
                    p.queries.size(), (unsigned long long)p.quotient_num_chunk, (unsigned long long)p.quotient_chunk_degree,
                    p.quotient_numerator.fields.empty() ? "" : p.quotient_numerator.fields[0].first.c_str());
        int64_t min_rot = 0, max_rot = 0;
        for (auto& q : p.queries) { min_rot = std::min(min_rot, q.rotation); max_rot = std::max(max_rot, q.rotation); }
        std::printf("\"rotations\": [%lld, %lld], \"has_initial_state\": %s, ", (long long)min_rot, (long long)max_rot, p.has_transcript_initial_state ? "true" : "false");
        if (p.has_transcript_initial_state) print_limbs("transcript_initial_state", p.transcript_initial_state);
        std::printf("\"instances_committed\": %s, \"accumulator_limbs\": %zu, \"proof_bytes_shplonk\": %llu, \"vk_match\": %d}\n",
                    p.instances_committed ? "true" : "false", p.accumulator_indices.empty() ? 0 : p.accumulator_indices[0].size(),
                    (unsigned long long)protocol::proof_bytes_shplonk(p), vk_match);
        return 0;
    } catch (const std::exception& e) {
        std::printf("{\"error\": \"%s\"}\n", e.what());
        return 1;
    }
}
