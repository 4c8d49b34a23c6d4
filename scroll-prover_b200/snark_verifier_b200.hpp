// snark_verifier_b200.hpp — host-side verifier for the reference's Poseidon-transcript SHPLONK proofs, driven by a `*.protocol` file.
//
// C++ mirror (the reference is Rust) of snark-verifier's native verifier as scroll-prover uses it for chunk / batch proofs
// (`ChunkVerifier::verify_chunk_proof`, `BatchVerifier::verify_batch_proof`: /root/reference/integration/src/prove.rs:50-53,78-80;
// snark-verifier @ 948671c, /root/reference/Cargo.lock:3948-3950):
//   PlonkProof::read / PlonkVerifier::verify          (verifier/plonk.rs, verifier/plonk/protocol.rs: Expression, CommonPolynomial)
//   PoseidonTranscript<NativeLoader>                  (system/halo2/transcript/halo2.rs; sponge in plonk_b200.hpp)
//   Bdfg21 (SHPLONK) -> KzgAccumulator                (pcs/kzg/multiopen/bdfg21.rs, pcs/kzg/accumulator.rs)
//   KzgDecidingKey: e(lhs, g2) = e(rhs, s_g2)         (pcs/kzg/decider.rs; on chain: precompile 0x08, evm_verifier.yul:1240)
// Inputs are the artefacts the reference ships: the protocol JSON (protocol_json.hpp), the instance values, the proof bytes and the
// SRS's two G2 points.  Host-only: verification is not on the hot path; this exists so that proofs -- the reference's own, and the
// ones made through the C ABI by plonk_b200.hpp -- can be checked without the Rust stack.  tests/test_snark_verifier_host.py runs it on
// the reference's shipped proofs next to the independent big-integer model (tests/snark_verifier_model.py).
#pragma once
#include "plonk_b200.hpp"
#include "protocol_json.hpp"

namespace halo2_b200 {
namespace snark {

using plonk::f_add;
using plonk::f_inv;
using plonk::f_mul;
using plonk::f_neg;
using plonk::f_one;
using plonk::f_pow;
using plonk::f_sub;
using plonk::f_zero;

struct KzgAccumulator {
    pairing::G1Point lhs, rhs;
};

inline Fr fr_of(const protocol::Limbs4& l) {  // the protocol stores raw Montgomery limbs: our Fr as is
    Fr r;
    std::memcpy(r.l, l.l, 32);
    return r;
}
inline serde::G1Point point_of(const protocol::Point& p) {
    serde::G1Point r;
    std::memcpy(r.x.l.v, p.x.l, 32);
    std::memcpy(r.y.l.v, p.y.l, 32);
    return r;
}

// PlonkVerifier::verify up to the accumulator: throws Panic on malformed proofs (bad encodings, wrong length)
inline KzgAccumulator verify_to_accumulator(const protocol::PlonkProtocol& P, const std::vector<std::vector<Fr>>& instances,
                                            const std::vector<uint8_t>& proof) {
    using plonk::hostg1::XYZZ;
    if (P.instances_committed) throw Panic("snark verifier: committed instances are not supported");
    if (instances.size() != P.num_instance.size()) throw Panic("snark verifier: wrong number of instance columns");
    for (size_t i = 0; i < instances.size(); ++i)
        if (instances[i].size() != P.num_instance[i]) throw Panic("snark verifier: wrong number of instances");
    const Fr omega = fr_of(P.domain.gen), omega_inv = fr_of(P.domain.gen_inv), n_inv = fr_of(P.domain.n_inv);
    const uint64_t n = P.domain.n;
    plonk::Transcript tr(proof, plonk::TranscriptKind::Poseidon);
    if (P.has_transcript_initial_state) tr.common_scalar(fr_of(P.transcript_initial_state));
    for (auto& col : instances)
        for (auto& v : col) tr.common_scalar(v);
    std::vector<serde::G1Point> witnesses, quotients;
    std::vector<Fr> challenges;
    for (size_t ph = 0; ph < P.num_witness.size(); ++ph) {
        for (uint64_t i = 0; i < P.num_witness[ph]; ++i) witnesses.push_back(tr.read_point());
        for (uint64_t i = 0; i < P.num_challenge[ph]; ++i) challenges.push_back(tr.squeeze_challenge());
    }
    for (uint64_t i = 0; i < P.quotient_num_chunk; ++i) quotients.push_back(tr.read_point());
    const Fr z = tr.squeeze_challenge();
    std::vector<Fr> evaluations;
    for (size_t i = 0; i < P.evaluations.size(); ++i) evaluations.push_back(tr.read_scalar());
    const Fr mu = tr.squeeze_challenge(), gamma = tr.squeeze_challenge();
    const serde::G1Point w = tr.read_point();
    const Fr z_prime = tr.squeeze_challenge();
    const serde::G1Point w_prime = tr.read_point();
    if (!tr.exhausted()) throw Panic("snark verifier: trailing bytes after the proof");

    const Fr zn = f_pow(z, n);
    auto omega_pow = [&](int64_t i) { return i >= 0 ? f_pow(omega, (uint64_t)i) : f_pow(omega_inv, (uint64_t)(-i)); };
    auto lagrange = [&](int64_t i) {
        Fr wi = omega_pow(i);
        return f_mul(f_mul(f_mul(wi, n_inv), f_sub(zn, f_one())), f_inv(f_sub(z, wi)));
    };
    const size_t n_pre = P.preprocessed.size(), n_inst = instances.size();
    const size_t quotient_poly = n_pre + n_inst + witnesses.size();
    Fr quotient_eval = f_zero();
    bool have_quotient = false;
    auto poly_eval = [&](uint64_t poly, int64_t rot) -> Fr {
        if (poly == quotient_poly && rot == 0 && have_quotient) return quotient_eval;
        for (size_t i = 0; i < P.evaluations.size(); ++i)
            if (P.evaluations[i].poly == poly && P.evaluations[i].rotation == rot) return evaluations[i];
        if (poly >= n_pre && poly < n_pre + n_inst) {  // instances are not committed: interpolate
            Fr acc = f_zero();
            const auto& col = instances[poly - n_pre];
            for (size_t i = 0; i < col.size(); ++i)
                if (!plonk::f_is_zero(col[i])) acc = f_add(acc, f_mul(col[i], lagrange((int64_t)i - rot)));
            return acc;
        }
        throw Panic("snark verifier: the protocol queries a polynomial at a point that was not opened");
    };
    std::function<Fr(const protocol::Json&)> ev = [&](const protocol::Json& e) -> Fr {
        if (e.kind != protocol::Json::Object || e.fields.size() != 1) throw Panic("snark verifier: malformed expression node");
        const std::string& kind = e.fields[0].first;
        const protocol::Json& a = e.fields[0].second;
        if (kind == "Constant") return fr_of(protocol::limbs(a));
        if (kind == "CommonPolynomial") {
            if (a.kind == protocol::Json::String) {
                if (a.text == "Identity") return z;
                throw Panic("snark verifier: unknown common polynomial " + a.text);
            }
            return lagrange(a.at("Lagrange").i64());
        }
        if (kind == "Polynomial") return poly_eval(a.at("poly").u64(), a.at("rotation").i64());
        if (kind == "Challenge") return challenges.at((size_t)a.u64());
        if (kind == "Negated") return f_neg(ev(a));
        if (kind == "Sum") return f_add(ev(a.items.at(0)), ev(a.items.at(1)));
        if (kind == "Product") return f_mul(ev(a.items.at(0)), ev(a.items.at(1)));
        if (kind == "Scaled") return f_mul(ev(a.items.at(0)), fr_of(protocol::limbs(a.items.at(1))));
        if (kind == "DistributePowers") {
            const auto& exprs = a.items.at(0).items;
            const Fr base = ev(a.items.at(1));
            Fr acc = ev(exprs.at(0));
            for (size_t i = 1; i < exprs.size(); ++i) acc = f_add(f_mul(acc, base), ev(exprs[i]));
            return acc;
        }
        throw Panic("snark verifier: unknown expression node " + kind);
    };
    quotient_eval = f_mul(ev(P.quotient_numerator), f_inv(f_sub(zn, f_one())));
    have_quotient = true;

    // commitments by polynomial index; the quotient's is sum_i (z^n)^(chunk_degree * i) [Q_i]
    std::vector<XYZZ> commitments;
    for (auto& p : P.preprocessed) commitments.push_back(plonk::hostg1::from_point(point_of(p)));
    for (size_t i = 0; i < n_inst; ++i) commitments.push_back(XYZZ::identity());
    for (auto& p : witnesses) commitments.push_back(plonk::hostg1::from_point(p));
    {
        const Fr step = f_pow(zn, P.quotient_chunk_degree);
        XYZZ acc = XYZZ::identity();
        Fr pw = f_one();
        for (auto& q : quotients) {
            XYZZ t = plonk::hostg1::mul(q, pw);
            b200zk::xyzz_add(acc, t);
            pw = f_mul(pw, step);
        }
        commitments.push_back(acc);
    }
    // ---- Bdfg21: query sets in first-appearance order
    struct PolyShifts { uint64_t poly; std::vector<Fr> shifts, evals; };
    std::vector<PolyShifts> per_poly;
    for (auto& q : P.queries) {
        const Fr shift = omega_pow(q.rotation), e = poly_eval(q.poly, q.rotation);
        PolyShifts* ps = nullptr;
        for (auto& x : per_poly)
            if (x.poly == q.poly) ps = &x;
        if (!ps) {
            per_poly.push_back({q.poly, {}, {}});
            ps = &per_poly.back();
        }
        bool seen = false;
        for (auto& s : ps->shifts) seen |= (s == shift);
        if (!seen) {
            ps->shifts.push_back(shift);
            ps->evals.push_back(e);
        }
    }
    struct Set { std::vector<Fr> shifts; std::vector<uint64_t> polys; std::vector<std::vector<Fr>> evals; };
    std::vector<Set> sets;
    for (auto& pp : per_poly) {
        Set* target = nullptr;
        for (auto& s : sets) {
            if (s.shifts.size() != pp.shifts.size()) continue;
            bool all = true;
            for (auto& a : pp.shifts) {
                bool in = false;
                for (auto& b : s.shifts) in |= (a == b);
                all &= in;
            }
            if (all) { target = &s; break; }
        }
        if (!target) {
            sets.push_back({pp.shifts, {}, {}});
            target = &sets.back();
        }
        std::vector<Fr> evs;
        for (auto& sh : target->shifts)
            for (size_t i = 0; i < pp.shifts.size(); ++i)
                if (pp.shifts[i] == sh) evs.push_back(pp.evals[i]);
        target->polys.push_back(pp.poly);
        target->evals.push_back(evs);
    }
    std::vector<Fr> z_s(sets.size());
    for (size_t i = 0; i < sets.size(); ++i) {
        Fr v = f_one();
        for (auto& sh : sets[i].shifts) v = f_mul(v, f_sub(z_prime, f_mul(sh, z)));
        z_s[i] = v;
    }
    XYZZ f = XYZZ::identity();
    Fr constant = f_zero(), gpow = f_one();
    for (size_t i = 0; i < sets.size(); ++i) {
        const Fr coeff = f_mul(gpow, f_mul(z_s[0], f_inv(z_s[i])));
        std::vector<Fr> pts;
        for (auto& sh : sets[i].shifts) pts.push_back(f_mul(sh, z));
        Fr mpow = f_one();
        for (size_t j = 0; j < sets[i].polys.size(); ++j) {
            const Fr r_eval = plonk::eval_small(plonk::lagrange_interpolate(pts, sets[i].evals[j]), z_prime);
            const Fr c = f_mul(coeff, mpow);
            pairing::G1Point cp = plonk::hostg1::to_pairing_point(commitments.at(sets[i].polys[j]));
            XYZZ t = plonk::hostg1::mul(serde::G1Point{cp.x, cp.y}, c);
            b200zk::xyzz_add(f, t);
            constant = f_add(constant, f_mul(c, r_eval));
            mpow = f_mul(mpow, mu);
        }
        gpow = f_mul(gpow, gamma);
    }
    serde::G1Point gen;
    gen.x = b200zk::Fq::one();
    gen.y = b200zk::Fq::one().dbl();
    XYZZ t = plonk::hostg1::mul(gen, f_neg(constant));
    b200zk::xyzz_add(f, t);
    t = plonk::hostg1::mul(w, f_neg(z_s[0]));
    b200zk::xyzz_add(f, t);
    t = plonk::hostg1::mul(w_prime, z_prime);
    b200zk::xyzz_add(f, t);  // lhs = f + z' W'
    return {plonk::hostg1::to_pairing_point(f), pairing::G1Point{w_prime.x, w_prime.y}};
}

// KzgDecidingKey::decide for one accumulator: e(lhs, g2) = e(rhs, s_g2)
inline bool decide(const KzgAccumulator& acc, const pairing::G2Point& g2, const pairing::G2Point& s_g2) {
    if (!pairing::g1_on_curve(acc.lhs) || !pairing::g1_on_curve(acc.rhs)) return false;
    return pairing::pairing_check({{acc.lhs, g2}, {pairing::G1Point{acc.rhs.x, acc.rhs.y.neg()}, s_g2}});
}

// the accumulator an aggregation / compression proof carries in its instances at `accumulator_indices` (limbs of 88 bits, 3 per
// coordinate: lhs.x, lhs.y, rhs.x, rhs.y); false if the protocol carries none or a coordinate is not a field element
inline bool carried_accumulator(const protocol::PlonkProtocol& P, const std::vector<std::vector<Fr>>& instances, KzgAccumulator* out) {
    if (P.accumulator_indices.empty() || P.accumulator_indices[0].size() != 12) return false;
    uint8_t be[384];
    for (int i = 0; i < 12; ++i) {
        auto idx = P.accumulator_indices[0][i];
        uint8_t le[32];
        plonk::f_to_repr(instances.at(idx.first).at(idx.second), le);
        for (int b = 0; b < 32; ++b) be[32 * i + b] = le[31 - b];
    }
    return pairing::accumulator_from_limbs(be, &out->lhs, &out->rhs);
}

// verify_chunk_proof / verify_batch_proof: the proof's own accumulator and the one it carries forward must both be valid
inline bool verify(const protocol::PlonkProtocol& P, const std::vector<std::vector<Fr>>& instances, const std::vector<uint8_t>& proof,
                   const pairing::G2Point& g2, const pairing::G2Point& s_g2, std::string* why = nullptr) {
    try {
        KzgAccumulator acc = verify_to_accumulator(P, instances, proof);
        if (!decide(acc, g2, s_g2)) {
            if (why) *why = "the proof's accumulator is not valid";
            return false;
        }
        KzgAccumulator carried;
        if (!P.accumulator_indices.empty()) {
            if (!carried_accumulator(P, instances, &carried) || !decide(carried, g2, s_g2)) {
                if (why) *why = "the accumulator carried in the instances is not valid";
                return false;
            }
        }
        return true;
    } catch (const std::exception& e) {
        if (why) *why = e.what();
        return false;
    }
}

}  // namespace snark
}  // namespace halo2_b200
