// proof_files.hpp — reader for the proof files the reference writes and consumes (on-disk formats):
//   * a chunk / batch proof object as `prover` serialises it with serde_json (what `gen_halo2_chunk_proof` / `gen_batch_proof` return
//     and `verify_chunk_proof` / `verify_batch_proof` take, /root/reference/integration/src/prove.rs:37-39,50-53,69-80):
//       {"protocol": base64(<PlonkProtocol JSON, see protocol_json.hpp>), "proof": base64(<proof bytes>),
//        "instances": base64(<32-byte BIG-endian words of the single instance column>), "vk": base64(<vk_*.vkey bytes>),
//        "git_version": "...", ("chunk_info": {...}, "row_usages": [...]) | ("batch_hash": "0x...")}
//     e.g. /root/reference/integration/tests/test_data/full_proof_batch_agg_1.json;
//   * the containers around them: `{"chunk_proofs": [<chunk proof>, ...], ...}` — a batch proving task
//     (test_data/batch-task-*.json, batch_tasks/*.json) or a dumped chunk-proof list (full_proof_1.json).
// With snark_verifier_b200.hpp this is `ChunkVerifier::verify_chunk_proof` / `BatchVerifier::verify_batch_proof` on the reference's own
// files: verify_entry() checks the proof under the protocol it carries, the accumulator it carries forward, and that the verifying
// key bytes beside it hold exactly the protocol's preprocessed commitments (same circuit).  Host-only, header-only.
#pragma once
#include <string>
#include <vector>

#include "snark_verifier_b200.hpp"

namespace halo2_b200 {
namespace proof_files {

// RFC 4648 base64 (standard alphabet, '=' padding; serde's base64 encoding of Vec<u8>); throws on a foreign character
inline std::vector<uint8_t> base64_decode(const std::string& s) {
    auto val = [](char c) -> int {
        if (c >= 'A' && c <= 'Z') return c - 'A';
        if (c >= 'a' && c <= 'z') return c - 'a' + 26;
        if (c >= '0' && c <= '9') return c - '0' + 52;
        if (c == '+') return 62;
        if (c == '/') return 63;
        return -1;
    };
    std::vector<uint8_t> out;
    out.reserve(s.size() / 4 * 3);
    uint32_t acc = 0;
    int bits = 0;
    size_t i = 0;
    for (; i < s.size() && s[i] != '='; ++i) {
        int v = val(s[i]);
        if (v < 0) throw std::runtime_error("proof file: not base64");
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            out.push_back((uint8_t)(acc >> bits));
            acc &= (1u << bits) - 1;
        }
    }
    for (size_t j = i; j < s.size(); ++j)
        if (s[j] != '=') throw std::runtime_error("proof file: data after base64 padding");
    if (acc != 0 || (s.size() - i) > 2 || s.size() % 4 != 0) throw std::runtime_error("proof file: malformed base64 tail");
    return out;
}

struct ProofEntry {
    protocol::PlonkProtocol protocol;
    std::vector<uint8_t> proof, vk;
    std::vector<std::vector<Fr>> instances;  // one column
    std::string git_version;
    bool is_chunk = false;  // carries a chunk_info (chunk proof) rather than a batch_hash (batch proof)
};

inline ProofEntry parse_entry(const protocol::Json& e) {
    ProofEntry p;
    const std::vector<uint8_t> proto = base64_decode(e.at("protocol").text);
    p.protocol = protocol::parse_protocol(std::string(proto.begin(), proto.end()));
    p.proof = base64_decode(e.at("proof").text);
    p.vk = base64_decode(e.at("vk").text);
    const std::vector<uint8_t> inst = base64_decode(e.at("instances").text);
    if (inst.size() % 32) throw std::runtime_error("proof file: instances are 32-byte words");
    std::vector<Fr> col;
    for (size_t off = 0; off < inst.size(); off += 32) {
        uint8_t le[32];
        for (int b = 0; b < 32; ++b) le[b] = inst[off + 31 - b];
        Fr v;
        if (!plonk::f_from_repr(le, &v)) throw std::runtime_error("proof file: an instance is not a canonical field element");
        col.push_back(v);
    }
    p.instances.push_back(col);
    if (e.has("git_version")) p.git_version = e.at("git_version").text;
    p.is_chunk = e.has("chunk_info");
    return p;
}

// every proof object of a file: the entries of "chunk_proofs" if the file is a container, else the file itself
inline std::vector<ProofEntry> parse_file(const std::string& text) {
    protocol::Json j = protocol::JsonParser(text).parse();
    std::vector<ProofEntry> out;
    if (j.has("chunk_proofs")) {
        for (auto& e : j.at("chunk_proofs").items) out.push_back(parse_entry(e));
    } else {
        out.push_back(parse_entry(j));
    }
    return out;
}

// the verifying key stored beside a proof is the one its protocol was compiled from: k and the fixed + permutation commitments
inline bool vk_matches_protocol(const ProofEntry& p, std::string* why = nullptr) {
    serde::VerifyingKeyFile vk;
    if (!serde::read_vk_processed(p.vk.data(), p.vk.size(), &vk)) {
        if (why) *why = "the vk bytes do not parse";
        return false;
    }
    std::vector<serde::G1Point> pts = vk.fixed_commitments;
    pts.insert(pts.end(), vk.permutation_commitments.begin(), vk.permutation_commitments.end());
    bool same = vk.k == p.protocol.domain.k && pts.size() == p.protocol.preprocessed.size();
    for (size_t i = 0; same && i < pts.size(); ++i) {
        const serde::G1Point q = snark::point_of(p.protocol.preprocessed[i]);
        same = pts[i].x == q.x && pts[i].y == q.y;
    }
    if (!same && why) *why = "the vk's commitments are not the protocol's preprocessed polynomials";
    return same;
}

// ChunkVerifier::verify_chunk_proof / BatchVerifier::verify_batch_proof on one proof object
inline bool verify_entry(const ProofEntry& p, const pairing::G2Point& g2, const pairing::G2Point& s_g2, std::string* why = nullptr) {
    if (!vk_matches_protocol(p, why)) return false;
    return snark::verify(p.protocol, p.instances, p.proof, g2, s_g2, why);
}

}  // namespace proof_files
}  // namespace halo2_b200
