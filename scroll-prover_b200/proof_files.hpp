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
// key bytes beside it hold exactly the protocol's preprocessed commitments (same circuit).  A chunk proof's public input is bound to
// its `chunk_info` as well: instances[12..44] are the 32 bytes of
//   keccak256(chain_id as u64 BE || prev_state_root || post_state_root || withdraw_root || data_hash || keccak256(tx_bytes))
// (`ChunkInfo::public_input_hash`, crate `aggregator` of scroll-tech/zkevm-circuits @ 7fd6b6d, /root/reference/Cargo.lock:32-34 -- an
// un-vendored dependency; the formula is anchored on the reference's data: it reproduces the public input of all 319 chunk proofs
// under integration/tests/test_data).  A batch proof's `batch_hash` field is checked against its public input likewise.
// Host-only, header-only.
#pragma once
#include <array>
#include <string>
#include <vector>

#include "keccak256.hpp"
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

using hash::keccak256;

struct ChunkInfo {  // the fields of a proof object's "chunk_info" that enter the public input
    uint64_t chain_id = 0;
    std::array<uint8_t, 32> prev_state_root{}, post_state_root{}, withdraw_root{}, data_hash{};
    std::vector<uint8_t> tx_bytes;
    std::array<uint8_t, 32> public_input_hash() const {
        std::vector<uint8_t> pre;
        for (int b = 7; b >= 0; --b) pre.push_back((uint8_t)(chain_id >> (8 * b)));
        for (auto* f : {&prev_state_root, &post_state_root, &withdraw_root, &data_hash}) pre.insert(pre.end(), f->begin(), f->end());
        const auto txh = keccak256(tx_bytes.data(), tx_bytes.size());
        pre.insert(pre.end(), txh.begin(), txh.end());
        return keccak256(pre.data(), pre.size());
    }
};

inline std::array<uint8_t, 32> hex32(const std::string& h0) {  // 64 hex digits, with or without "0x" (the reference's files hold both)
    const std::string h = h0.size() == 64 ? "0x" + h0 : h0;
    if (h.size() != 66 || h[0] != '0' || h[1] != 'x') throw std::runtime_error("proof file: expected a 32-byte hex string");
    std::array<uint8_t, 32> out;
    auto nib = [](char c) -> int { return c >= '0' && c <= '9' ? c - '0' : (c >= 'a' && c <= 'f' ? c - 'a' + 10 : (c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1)); };
    for (int i = 0; i < 32; ++i) {
        int hi = nib(h[2 + 2 * i]), lo = nib(h[3 + 2 * i]);
        if (hi < 0 || lo < 0) throw std::runtime_error("proof file: not a hex digit");
        out[i] = (uint8_t)(hi * 16 + lo);
    }
    return out;
}

inline ChunkInfo parse_chunk_info(const protocol::Json& ci) {
    ChunkInfo c;
    c.chain_id = ci.at("chain_id").u64();
    c.prev_state_root = hex32(ci.at("prev_state_root").text);
    c.post_state_root = hex32(ci.at("post_state_root").text);
    c.withdraw_root = hex32(ci.at("withdraw_root").text);
    c.data_hash = hex32(ci.at("data_hash").text);
    c.tx_bytes = base64_decode(ci.at("tx_bytes").text);
    return c;
}

// the "batch_header" of a batch proving task.  batch_hash = keccak256 of its 193-byte encoding
//   version (1) | batch_index | l1_message_popped | total_l1_message_popped (8 each, BE) | data_hash | blob_versioned_hash |
//   parent_batch_hash (32 each) | last_block_timestamp (8, BE) | blob_data_proof (2 x 32)
// (`BatchHeader::batch_hash`, crate `aggregator`, same pin as above).  Anchored on the reference's data: the header of
// test_data/full_proof_batch_prove_1.json hashes to the `batch_hash` of full_proof_batch_agg_1.json, and the ten consecutive tasks
// batch_tasks/batch_task_2932{05..14}.json chain through `parent_batch_hash`; `data_hash` = keccak256 of the chunks' data hashes.
struct BatchHeader {
    uint8_t version = 0;
    uint64_t batch_index = 0, l1_message_popped = 0, total_l1_message_popped = 0, last_block_timestamp = 0;
    std::array<uint8_t, 32> data_hash{}, blob_versioned_hash{}, parent_batch_hash{}, blob_data_proof[2] = {{}, {}};
    std::vector<uint8_t> encode() const {
        std::vector<uint8_t> out;
        auto be64 = [&](uint64_t v) { for (int b = 7; b >= 0; --b) out.push_back((uint8_t)(v >> (8 * b))); };
        auto put = [&](const std::array<uint8_t, 32>& a) { out.insert(out.end(), a.begin(), a.end()); };
        out.push_back(version);
        be64(batch_index); be64(l1_message_popped); be64(total_l1_message_popped);
        put(data_hash); put(blob_versioned_hash); put(parent_batch_hash);
        be64(last_block_timestamp);
        put(blob_data_proof[0]); put(blob_data_proof[1]);
        return out;
    }
    std::array<uint8_t, 32> batch_hash() const {
        const std::vector<uint8_t> e = encode();
        return keccak256(e.data(), e.size());
    }
};
inline BatchHeader parse_batch_header(const protocol::Json& h) {
    BatchHeader b;
    b.version = (uint8_t)h.at("version").u64();
    b.batch_index = h.at("batch_index").u64();
    b.l1_message_popped = h.at("l1_message_popped").u64();
    b.total_l1_message_popped = h.at("total_l1_message_popped").u64();
    b.last_block_timestamp = h.at("last_block_timestamp").u64();
    b.data_hash = hex32(h.at("data_hash").text);
    b.blob_versioned_hash = hex32(h.at("blob_versioned_hash").text);
    b.parent_batch_hash = hex32(h.at("parent_batch_hash").text);
    if (h.at("blob_data_proof").items.size() != 2) throw std::runtime_error("proof file: blob_data_proof holds two words");
    for (int i = 0; i < 2; ++i) b.blob_data_proof[i] = hex32(h.at("blob_data_proof").items[i].text);
    return b;
}

struct ProofEntry {
    protocol::PlonkProtocol protocol;
    std::vector<uint8_t> proof, vk;
    std::vector<std::vector<Fr>> instances;  // one column
    std::string git_version;
    bool is_chunk = false;  // carries a chunk_info (chunk proof) rather than a batch_hash (batch proof)
    ChunkInfo chunk_info;
    bool has_batch_hash = false;
    std::array<uint8_t, 32> batch_hash{};
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
    if (p.is_chunk) p.chunk_info = parse_chunk_info(e.at("chunk_info"));
    if (e.has("batch_hash")) {
        p.has_batch_hash = true;
        p.batch_hash = hex32(e.at("batch_hash").text);
    }
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

// a batch proving task (`BatchProvingTask`: what `gen_batch_proof` takes, /root/reference/integration/src/prove.rs:69): the chunk
// proofs to aggregate, the chunk infos they were made for, the header of the batch
struct BatchTask {
    std::vector<ChunkInfo> chunk_infos;
    std::vector<ProofEntry> chunk_proofs;
    BatchHeader header;
};
inline BatchTask parse_batch_task(const std::string& text) {
    protocol::Json j = protocol::JsonParser(text).parse();
    BatchTask t;
    for (auto& c : j.at("chunk_infos").items) t.chunk_infos.push_back(parse_chunk_info(c));
    for (auto& e : j.at("chunk_proofs").items) t.chunk_proofs.push_back(parse_entry(e));
    t.header = parse_batch_header(j.at("batch_header"));
    return t;
}
// the consistency a batch prover relies on before aggregating: one chunk info per proof and the same one the proof carries, the state
// roots of consecutive chunks chained, the header's data_hash = keccak256(chunk data hashes)
inline bool check_batch_task(const BatchTask& t, std::string* why = nullptr) {
    auto fail = [&](const char* m) { if (why) *why = m; return false; };
    if (t.chunk_infos.empty() || t.chunk_infos.size() != t.chunk_proofs.size()) return fail("one chunk info per chunk proof");
    std::vector<uint8_t> hashes;
    for (size_t i = 0; i < t.chunk_infos.size(); ++i) {
        if (!t.chunk_proofs[i].is_chunk || t.chunk_infos[i].public_input_hash() != t.chunk_proofs[i].chunk_info.public_input_hash())
            return fail("a chunk info differs from the one its proof carries");
        if (i && t.chunk_infos[i].prev_state_root != t.chunk_infos[i - 1].post_state_root) return fail("the chunks' state roots do not chain");
        hashes.insert(hashes.end(), t.chunk_infos[i].data_hash.begin(), t.chunk_infos[i].data_hash.end());
    }
    if (keccak256(hashes.data(), hashes.size()) != t.header.data_hash) return fail("the header's data_hash is not the hash of the chunks' data hashes");
    return true;
}

// the public input of the batch proof made from a task, after the 12 accumulator limbs: (high, low) 16-byte halves of the first chunk's
// prev_state_root, the header's parent_batch_hash, the last chunk's post_state_root, the header's batch hash; the chain id; the
// halves of the last chunk's withdraw_root.  Read off and checked on test_data/full_proof_batch_prove_1.json (task) and
// full_proof_batch_agg_1.json (the batch proof the reference made from it).
inline std::vector<Fr> batch_public_input(const BatchTask& t) {
    if (t.chunk_infos.empty()) throw std::runtime_error("proof file: a batch without chunks");
    std::vector<Fr> out;
    auto halves = [&](const std::array<uint8_t, 32>& v) {
        for (int half = 0; half < 2; ++half) {
            uint8_t le[32] = {0};
            for (int b = 0; b < 16; ++b) le[b] = v[16 * half + 15 - b];
            Fr f;
            plonk::f_from_repr(le, &f);
            out.push_back(f);
        }
    };
    halves(t.chunk_infos.front().prev_state_root);
    halves(t.header.parent_batch_hash);
    halves(t.chunk_infos.back().post_state_root);
    halves(t.header.batch_hash());
    out.push_back(plonk::f_u64(t.chunk_infos.front().chain_id));
    halves(t.chunk_infos.back().withdraw_root);
    return out;
}
inline bool batch_proof_matches_task(const ProofEntry& batch_proof, const BatchTask& t) {
    const std::vector<Fr> pi = batch_public_input(t);
    if (batch_proof.instances.size() != 1 || batch_proof.instances[0].size() != 12 + pi.size()) return false;
    for (size_t i = 0; i < pi.size(); ++i)
        if (!(batch_proof.instances[0][12 + i] == pi[i])) return false;
    return true;
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

// a chunk proof's public input: 12 accumulator limbs, then the 32 bytes of ChunkInfo::public_input_hash, one per instance cell
inline bool chunk_info_matches_instances(const ProofEntry& p, std::string* why = nullptr) {
    const auto h = p.chunk_info.public_input_hash();
    bool same = p.is_chunk && p.instances.size() == 1 && p.instances[0].size() == 12 + 32;
    for (int i = 0; same && i < 32; ++i) same = p.instances[0][12 + i] == plonk::f_u64(h[i]);
    if (!same && why) *why = "the public input is not the hash of the chunk_info beside the proof";
    return same;
}

// a batch proof's public input: 12 accumulator limbs, then 32-byte values as (high, low) 16-byte halves -- parent_state_root,
// parent_batch_hash, current_state_root, batch_hash -- then chain_id and withdraw_root (high, low); the "batch_hash" field beside the
// proof is cells 18 and 19 (layout read off the reference's two shipped batch proofs, whose second continues the first)
inline bool batch_hash_matches_instances(const ProofEntry& p, std::string* why = nullptr) {
    bool same = p.has_batch_hash && p.instances.size() == 1 && p.instances[0].size() == 12 + 11;
    for (int half = 0; same && half < 2; ++half) {
        uint8_t le[32] = {0};
        for (int b = 0; b < 16; ++b) le[b] = p.batch_hash[16 * half + 15 - b];
        Fr v;
        same = plonk::f_from_repr(le, &v) && p.instances[0][18 + half] == v;
    }
    if (!same && why) *why = "the public input does not carry the batch_hash beside the proof";
    return same;
}

// ChunkVerifier::verify_chunk_proof / BatchVerifier::verify_batch_proof on one proof object
inline bool verify_entry(const ProofEntry& p, const pairing::G2Point& g2, const pairing::G2Point& s_g2, std::string* why = nullptr) {
    if (!vk_matches_protocol(p, why)) return false;
    if (p.is_chunk && !chunk_info_matches_instances(p, why)) return false;
    if (p.has_batch_hash && !batch_hash_matches_instances(p, why)) return false;
    return snark::verify(p.protocol, p.instances, p.proof, g2, s_g2, why);
}

}  // namespace proof_files
}  // namespace halo2_b200
