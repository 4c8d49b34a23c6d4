// protocol_json.hpp — reader for the `*.protocol` files the reference ships beside its verifying keys
// (/root/reference/release-v0.13.1/chunk.protocol, integration/tests/test_data/chunk_chunk_0.protocol): snark-verifier's
// `PlonkProtocol<G1Affine>` serialised with serde_json -- the evaluation domain as raw Montgomery limbs, the preprocessed
// (fixed / permutation) commitments as affine points in Montgomery limbs, the shape of the proof (witness commitments and
// challenges per phase, evaluations, opening queries, quotient chunking), the transcript's initial state and the accumulator
// indices.  SURVEY.md §8(f).3 (on-disk formats).  Host-only, header-only, no dependencies: a ~100-line JSON value parser that
// keeps integers exact (the limbs are u64), and a typed view of the fields tooling needs -- e.g. to check a freshly built
// proving key against the shipped protocol: `domain` against EvaluationDomain::new_, `preprocessed` against the `vk_*.vkey`
// commitments (serde_bn254.hpp) or against device commitments of the fixed columns.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace halo2_b200 {
namespace protocol {

struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    std::string text;  // Number (verbatim digits: exact for u64 and beyond) or String
    std::vector<Json> items;
    std::vector<std::pair<std::string, Json>> fields;

    const Json& at(const std::string& key) const {
        for (auto& f : fields)
            if (f.first == key) return f.second;
        throw std::runtime_error("protocol json: missing field '" + key + "'");
    }
    bool has(const std::string& key) const {
        for (auto& f : fields)
            if (f.first == key) return true;
        return false;
    }
    uint64_t u64() const {
        if (kind != Number || text.empty() || text[0] == '-') throw std::runtime_error("protocol json: expected an unsigned integer");
        uint64_t v = 0;
        for (char c : text) {
            if (c < '0' || c > '9') throw std::runtime_error("protocol json: not an integer: " + text);
            uint64_t nv = v * 10 + (uint64_t)(c - '0');
            if (nv / 10 != v) throw std::runtime_error("protocol json: integer exceeds 64 bits: " + text);
            v = nv;
        }
        return v;
    }
    int64_t i64() const {
        if (kind != Number) throw std::runtime_error("protocol json: expected an integer");
        bool neg = !text.empty() && text[0] == '-';
        Json t = *this;
        if (neg) t.text = text.substr(1);
        uint64_t v = t.u64();
        return neg ? -(int64_t)v : (int64_t)v;
    }
};

class JsonParser {
  public:
    explicit JsonParser(const std::string& s) : s_(s) {}
    Json parse() {
        Json v = value();
        ws();
        if (p_ != s_.size()) fail("trailing characters");
        return v;
    }

  private:
    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("protocol json: ") + what + " at offset " + std::to_string(p_)); }
    void ws() {
        while (p_ < s_.size() && (s_[p_] == ' ' || s_[p_] == '\n' || s_[p_] == '\t' || s_[p_] == '\r')) ++p_;
    }
    bool lit(const char* w) {
        size_t n = std::char_traits<char>::length(w);
        if (s_.compare(p_, n, w) == 0) {
            p_ += n;
            return true;
        }
        return false;
    }
    std::string str() {
        std::string out;
        ++p_;  // opening quote
        while (p_ < s_.size() && s_[p_] != '"') {
            if (s_[p_] == '\\') {
                if (++p_ >= s_.size()) fail("bad escape");
                char c = s_[p_];
                out.push_back(c == 'n' ? '\n' : c == 't' ? '\t' : c);  // the protocol files hold no exotic escapes
            } else {
                out.push_back(s_[p_]);
            }
            ++p_;
        }
        if (p_ >= s_.size()) fail("unterminated string");
        ++p_;
        return out;
    }
    Json value() {
        ws();
        if (p_ >= s_.size()) fail("unexpected end");
        Json v;
        char c = s_[p_];
        if (c == '{') {
            v.kind = Json::Object;
            ++p_;
            ws();
            if (s_[p_] == '}') { ++p_; return v; }
            for (;;) {
                ws();
                if (s_[p_] != '"') fail("expected a field name");
                std::string k = str();
                ws();
                if (s_[p_++] != ':') fail("expected ':'");
                v.fields.emplace_back(k, value());
                ws();
                if (s_[p_] == ',') { ++p_; continue; }
                if (s_[p_] == '}') { ++p_; return v; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v.kind = Json::Array;
            ++p_;
            ws();
            if (s_[p_] == ']') { ++p_; return v; }
            for (;;) {
                v.items.push_back(value());
                ws();
                if (s_[p_] == ',') { ++p_; continue; }
                if (s_[p_] == ']') { ++p_; return v; }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v.kind = Json::String;
            v.text = str();
            return v;
        }
        if (lit("null")) return v;
        if (lit("true")) { v.kind = Json::Bool; v.b = true; return v; }
        if (lit("false")) { v.kind = Json::Bool; return v; }
        size_t q = p_;
        if (s_[q] == '-') ++q;
        while (q < s_.size() && ((s_[q] >= '0' && s_[q] <= '9') || s_[q] == '.' || s_[q] == 'e' || s_[q] == 'E' || s_[q] == '+' || s_[q] == '-')) ++q;
        if (q == p_) fail("unexpected character");
        v.kind = Json::Number;
        v.text = s_.substr(p_, q - p_);
        p_ = q;
        return v;
    }
    const std::string& s_;
    size_t p_ = 0;
};

struct Limbs4 {  // a field element exactly as serialised: 4 x u64 little-endian Montgomery limbs (halo2curves' raw repr)
    uint64_t l[4];
};
struct Point {
    Limbs4 x, y;
};
struct Query {
    uint64_t poly;
    int64_t rotation;
};
struct PlonkProtocol {
    struct Domain {
        uint32_t k;
        uint64_t n;
        Limbs4 n_inv, gen, gen_inv;
    } domain;
    std::vector<Point> preprocessed;
    std::vector<uint64_t> num_instance, num_witness, num_challenge;
    std::vector<Query> evaluations, queries;
    uint64_t quotient_num_chunk = 0, quotient_chunk_degree = 0;
    bool has_transcript_initial_state = false;
    Limbs4 transcript_initial_state{};
    bool instances_committed = false;  // instance_committing_key != null
    std::vector<std::vector<std::pair<uint64_t, uint64_t>>> accumulator_indices;
    Json quotient_numerator;  // the expression tree, kept as parsed JSON (Constant / CommonPolynomial / Polynomial / Challenge /
                              // Negated / Sum / Product / Scaled / DistributePowers nodes)
    // polynomial index layout of snark-verifier: preprocessed | instances | witnesses (all phases) | quotient
    uint64_t num_polys_before_quotient() const {
        uint64_t w = 0;
        for (auto v : num_witness) w += v;
        return preprocessed.size() + num_instance.size() + w;
    }
};

inline Limbs4 limbs(const Json& a) {
    if (a.kind != Json::Array || a.items.size() != 4) throw std::runtime_error("protocol json: a field element is 4 limbs");
    Limbs4 r;
    for (int i = 0; i < 4; ++i) r.l[i] = a.items[i].u64();
    return r;
}
inline std::vector<uint64_t> u64s(const Json& a) {
    std::vector<uint64_t> r;
    for (auto& v : a.items) r.push_back(v.u64());
    return r;
}
inline std::vector<Query> query_list(const Json& a) {
    std::vector<Query> r;
    for (auto& v : a.items) r.push_back({v.at("poly").u64(), v.at("rotation").i64()});
    return r;
}

inline PlonkProtocol parse_protocol(const std::string& text) {
    Json j = JsonParser(text).parse();
    PlonkProtocol p;
    const Json& d = j.at("domain");
    p.domain.k = (uint32_t)d.at("k").u64();
    p.domain.n = d.at("n").u64();
    if (p.domain.n != (1ull << p.domain.k)) throw std::runtime_error("protocol json: domain.n != 2^k");
    p.domain.n_inv = limbs(d.at("n_inv"));
    p.domain.gen = limbs(d.at("gen"));
    p.domain.gen_inv = limbs(d.at("gen_inv"));
    for (auto& pt : j.at("preprocessed").items) p.preprocessed.push_back({limbs(pt.at("x")), limbs(pt.at("y"))});
    p.num_instance = u64s(j.at("num_instance"));
    p.num_witness = u64s(j.at("num_witness"));
    p.num_challenge = u64s(j.at("num_challenge"));
    if (p.num_witness.size() != p.num_challenge.size()) throw std::runtime_error("protocol json: one challenge count per witness phase");
    p.evaluations = query_list(j.at("evaluations"));
    p.queries = query_list(j.at("queries"));
    const Json& q = j.at("quotient");
    p.quotient_num_chunk = q.at("num_chunk").u64();
    p.quotient_chunk_degree = q.at("chunk_degree").u64();
    p.quotient_numerator = q.at("numerator");
    const Json& tis = j.at("transcript_initial_state");
    if (tis.kind != Json::Null) {
        p.has_transcript_initial_state = true;
        p.transcript_initial_state = limbs(tis);
    }
    p.instances_committed = j.at("instance_committing_key").kind != Json::Null;
    for (auto& acc : j.at("accumulator_indices").items) {
        std::vector<std::pair<uint64_t, uint64_t>> one;
        for (auto& pr : acc.items) one.emplace_back(pr.items.at(0).u64(), pr.items.at(1).u64());
        p.accumulator_indices.push_back(one);
    }
    // every query names a polynomial that exists (the quotient polynomial is the one after the witnesses)
    for (auto* lst : {&p.evaluations, &p.queries})
        for (auto& qq : *lst)
            if (qq.poly > p.num_polys_before_quotient()) throw std::runtime_error("protocol json: query of an unknown polynomial");
    return p;
}

// size of a proof of this protocol with compressed 32-byte points and 32-byte scalars under SHPLONK (two opening points):
// witness commitments + quotient chunks + evaluations + 2
inline uint64_t proof_bytes_shplonk(const PlonkProtocol& p) {
    uint64_t w = 0;
    for (auto v : p.num_witness) w += v;
    return 32 * (w + p.quotient_num_chunk + p.evaluations.size() + 2);
}

}  // namespace protocol
}  // namespace halo2_b200
