// Driver for tests/test_snark_verifier_host.py: scroll-prover_b200/snark_verifier_b200.hpp on a case file
//   {"protocol": <protocol JSON object>, "proof": "<hex>", "instances": [["<32-byte big-endian hex>", ...], ...],
//    "s_g2": ["x_c1", "x_c0", "y_c1", "y_c0"]  (big-endian hex, EIP-197 word order; g2 is the generator)}
// prints ACCEPT or REJECT <reason>.
//   --file <proof file> <s_g2 x_c1> <x_c0> <y_c1> <y_c0>: a file AS THE REFERENCE WRITES IT (proof_files.hpp: a chunk / batch proof object
//   or a container with "chunk_proofs"); one line per proof object, `ACCEPT chunk|batch k=.. proof_bytes=.. git=..` or `REJECT <reason>`,
//   then `accepted <a> of <n>`; `--tamper` as a 7th argument flips one byte of every proof first.
//   --batch-task <file> [<batch proof file>]: a batch proving task (chunk_infos, chunk_proofs, batch_header): consistency + the header's batch hash.
//   --keccak <hex>: Keccak-256 of the message (proof_files.hpp's hash for the chunk public input).
#include <cstdio>
#include <fstream>
#include <sstream>

#include "../../scroll-prover_b200/proof_files.hpp"

using namespace halo2_b200;

static std::vector<uint8_t> unhex(const std::string& h) {
    std::vector<uint8_t> out;
    for (size_t i = 0; i + 1 < h.size(); i += 2) {
        unsigned v;
        std::sscanf(h.c_str() + i, "%2x", &v);
        out.push_back((uint8_t)v);
    }
    return out;
}

static int verify_file(int argc, char** argv) {
    try {
        std::ifstream f(argv[2], std::ios::binary);
        if (!f) {
            std::printf("REJECT cannot open %s\n", argv[2]);
            return 0;
        }
        std::stringstream ss;
        ss << f.rdbuf();
        uint8_t g2w[128];
        for (int i = 0; i < 4; ++i) {
            std::vector<uint8_t> w = unhex(argv[3 + i]);
            if (w.size() != 32) return 2;
            std::memcpy(g2w + 32 * i, w.data(), 32);
        }
        pairing::G2Point s_g2;
        if (!pairing::g2_from_eip197(g2w, &s_g2) || !pairing::g2_on_curve(s_g2)) {
            std::printf("REJECT malformed s_g2\n");
            return 0;
        }
        const bool tamper = argc > 7 && std::string(argv[7]) == "--tamper";
        std::vector<proof_files::ProofEntry> entries = proof_files::parse_file(ss.str());
        size_t accepted = 0;
        for (auto& e : entries) {
            if (tamper) e.proof[e.proof.size() / 3] ^= 0x10;
            std::string why;
            if (proof_files::verify_entry(e, pairing::g2_generator(), s_g2, &why)) {
                ++accepted;
                std::printf("ACCEPT %s k=%u proof_bytes=%zu git=%s\n", e.is_chunk ? "chunk" : "batch", e.protocol.domain.k, e.proof.size(), e.git_version.c_str());
            } else {
                std::printf("REJECT %s\n", why.c_str());
            }
        }
        std::printf("accepted %zu of %zu\n", accepted, entries.size());
        return 0;
    } catch (const std::exception& e) {
        std::printf("REJECT exception: %s\n", e.what());
        return 0;
    }
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (std::string(argv[1]) == "--file") return argc >= 7 ? verify_file(argc, argv) : 2;
    if (std::string(argv[1]) == "--batch-task" && argc >= 3) {  // --batch-task <file>: `CONSISTENT|INCONSISTENT <why>` batch_hash=.. parent=.. chunks=..
        try {
            std::ifstream f(argv[2], std::ios::binary);
            std::stringstream ss;
            ss << f.rdbuf();
            proof_files::BatchTask t = proof_files::parse_batch_task(ss.str());
            std::string why;
            const bool ok = proof_files::check_batch_task(t, &why);
            auto hex = [](const std::array<uint8_t, 32>& a) {
                std::string o;
                char b[3];
                for (uint8_t v : a) { std::snprintf(b, 3, "%02x", v); o += b; }
                return o;
            };
            std::printf("%s%s%s batch_hash=%s parent=%s chunks=%zu", ok ? "CONSISTENT" : "INCONSISTENT ", ok ? "" : why.c_str(), "", hex(t.header.batch_hash()).c_str(),
                        hex(t.header.parent_batch_hash).c_str(), t.chunk_proofs.size());
            if (argc >= 4) {  // a batch proof file: is its public input the one this task determines?
                std::ifstream g(argv[3], std::ios::binary);
                std::stringstream s2;
                s2 << g.rdbuf();
                std::printf(" proof_matches_task=%d", (int)proof_files::batch_proof_matches_task(proof_files::parse_file(s2.str()).at(0), t));
            }
            std::printf("\n");
        } catch (const std::exception& e) {
            std::printf("INCONSISTENT exception: %s\n", e.what());
        }
        return 0;
    }
    if (std::string(argv[1]) == "--keccak") {  // --keccak <hex message>: proof_files::keccak256, for the known-answer test
        std::vector<uint8_t> msg = unhex(argc > 2 ? argv[2] : "");
        auto d = proof_files::keccak256(msg.data(), msg.size());
        for (uint8_t b : d) std::printf("%02x", b);
        std::printf("\n");
        return 0;
    }
    try {
        std::ifstream f(argv[1], std::ios::binary);
        std::stringstream ss;
        ss << f.rdbuf();
        const std::string text = ss.str();
        protocol::Json j = protocol::JsonParser(text).parse();
        // re-serialising is not needed: the protocol object is parsed from its own span of the case file
        size_t p0 = text.find("\"protocol\"");
        p0 = text.find('{', p0);
        int depth = 0;
        size_t p1 = p0;
        for (; p1 < text.size(); ++p1) {
            if (text[p1] == '{') ++depth;
            if (text[p1] == '}' && --depth == 0) break;
        }
        protocol::PlonkProtocol P = protocol::parse_protocol(text.substr(p0, p1 - p0 + 1));
        std::vector<uint8_t> proof = unhex(j.at("proof").text);
        std::vector<std::vector<Fr>> instances;
        for (auto& col : j.at("instances").items) {
            std::vector<Fr> c;
            for (auto& w : col.items) {
                std::vector<uint8_t> be = unhex(w.text), le(32);
                for (int i = 0; i < 32; ++i) le[i] = be[31 - i];
                Fr v;
                if (!plonk::f_from_repr(le.data(), &v)) {
                    std::printf("REJECT instance is not a field element\n");
                    return 0;
                }
                c.push_back(v);
            }
            instances.push_back(c);
        }
        uint8_t g2w[128];
        for (int i = 0; i < 4; ++i) {
            std::vector<uint8_t> w = unhex(j.at("s_g2").items.at(i).text);
            std::memcpy(g2w + 32 * i, w.data(), 32);
        }
        pairing::G2Point s_g2;
        if (!pairing::g2_from_eip197(g2w, &s_g2) || !pairing::g2_on_curve(s_g2)) {
            std::printf("REJECT malformed s_g2\n");
            return 0;
        }
        std::string why;
        bool ok = snark::verify(P, instances, proof, pairing::g2_generator(), s_g2, &why);
        std::printf(ok ? "ACCEPT\n" : "REJECT %s\n", why.c_str());
        return 0;
    } catch (const std::exception& e) {
        std::printf("REJECT exception: %s\n", e.what());
        return 0;
    }
}
