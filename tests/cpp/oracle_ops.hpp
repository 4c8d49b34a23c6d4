// TEST INFRASTRUCTURE: halo2_b200::plonk::Ops over the CPU oracle (oracle/liboracle.so), so that the SAME create_proof code
// (scroll-prover_b200/plonk_b200.hpp) can be run once over the CUDA path and once over the restated reference arithmetic and
// the two proofs compared byte for byte.  Never part of the product.
#pragma once
#include "../../oracle/bn254_oracle.h"
#include "../../scroll-prover_b200/plonk_b200.hpp"

namespace oracle_ops {
using namespace halo2_b200;
using namespace halo2_b200::plonk;

class OracleOps : public Ops {
  public:
    OracleOps(const std::vector<G1Affine>& g, const std::vector<G1Affine>& g_lagrange, uint32_t j, uint32_t k) : g_(g), gl_(g_lagrange) {
        if (halo2_domain_new(&dom_, j, k) != 0) throw Panic("halo2_domain_new failed");
    }
    static const fr_t* fr(const Poly& p) { return reinterpret_cast<const fr_t*>(p.data()); }
    static fr_t* fr(Poly& p) { return reinterpret_cast<fr_t*>(p.data()); }
    static const fr_t* fr1(const Fr& v) { return reinterpret_cast<const fr_t*>(&v); }
    static G1 normalised(const g1_t& j) {  // same output convention as the ABI: (x, y, 1) or (0, 1, 0)
        G1 out;
        if (g1_is_identity(&j)) {
            std::memset(&out, 0, sizeof out);
            std::memcpy(out.y.l, fq_ONE.l, 32);
            return out;
        }
        g1_affine_t a;
        g1_to_affine(&a, &j);
        std::memcpy(out.x.l, a.x.l, 32);
        std::memcpy(out.y.l, a.y.l, 32);
        std::memcpy(out.z.l, fq_ONE.l, 32);
        return out;
    }
    G1 commit_lagrange(const Poly& v) override {
        g1_t r;
        halo2_commit(reinterpret_cast<const g1_affine_t*>(gl_.data()), fr(v), v.size(), 4, &r);
        return normalised(r);
    }
    G1 commit(const Poly& c) override {
        g1_t r;
        halo2_commit(reinterpret_cast<const g1_affine_t*>(g_.data()), fr(c), c.size(), 4, &r);
        return normalised(r);
    }
    Poly lagrange_to_coeff(Poly v) override {
        halo2_lagrange_to_coeff(&dom_, fr(v), 4);
        return v;
    }
    Poly coeff_to_extended(const Poly& c) override {
        Poly out((size_t)1 << dom_.extended_k);
        halo2_coeff_to_extended(&dom_, fr(c), fr(out), 4);
        return out;
    }
    Poly extended_to_coeff(Poly e) override {
        halo2_extended_to_coeff(&dom_, fr(e), 4);
        e.resize((size_t)dom_.n * dom_.quotient_poly_degree);
        return e;
    }
    Fr eval_polynomial(const Poly& c, const Fr& x) override {
        Fr r;
        halo2_eval_polynomial(reinterpret_cast<fr_t*>(&r), fr(c), c.size(), fr1(x));
        return r;
    }
    Poly kate_division(const Poly& c, const Fr& b) override {
        Poly q(c.size() - 1);
        halo2_kate_division(fr(q), fr(c), c.size(), fr1(b));
        return q;
    }
    Poly poly_mul(const Poly& a, const Poly& b) override {
        Poly r(a.size());
        for (size_t i = 0; i < a.size(); ++i) fr_mul(fr(r) + i, fr(a) + i, fr(b) + i);
        return r;
    }
    Poly poly_lincomb(const std::vector<const Poly*>& polys, const std::vector<Fr>& scalars) override {
        size_t n = 0;
        for (auto* p : polys) n = std::max(n, p->size());
        Poly r(n);
        std::memset(r.data(), 0, 32 * n);
        for (size_t j = 0; j < polys.size(); ++j)
            for (size_t i = 0; i < polys[j]->size(); ++i) {
                fr_t t;
                fr_mul(&t, fr(*polys[j]) + i, fr1(scalars[j]));
                fr_add(fr(r) + i, fr(r) + i, &t);
            }
        return r;
    }
    void graph_evaluate(const Program& p, const std::vector<const Poly*>& fixed, const std::vector<const Poly*>& advice,
                        const std::vector<const Poly*>& instance, const std::vector<Fr>& challenges, const Fr& beta, const Fr& gamma,
                        const Fr& theta, const Fr& y, Poly& values) override {
        auto tab = [](const std::vector<const Poly*>& v) {
            std::vector<const fr_t*> t;
            for (auto* c : v) t.push_back(fr(*c));
            return t;
        };
        auto tf = tab(fixed), ta = tab(advice), ti = tab(instance);
        static_assert(sizeof(halo2_calculation_t) == sizeof(b200zk_calculation) && sizeof(halo2_value_source_t) == sizeof(b200zk_value_source),
                      "the oracle's program structs mirror the ABI's");
        int rc = halo2_graph_evaluate(reinterpret_cast<const halo2_calculation_t*>(p.calcs.data()), (uint32_t)p.calcs.size(),
                                      reinterpret_cast<const halo2_value_source_t*>(p.parts.data()), fr(p.constants), p.rotations.data(),
                                      (uint32_t)p.rotations.size(), tf.data(), ta.data(), ti.data(),
                                      reinterpret_cast<const fr_t*>(challenges.data()), fr1(beta), fr1(gamma), fr1(theta), fr1(y),
                                      &dom_.extended_omega, fr(values), dom_.extended_k, 1 << (dom_.extended_k - dom_.k));
        if (rc != 0) throw Panic("halo2_graph_evaluate failed");
    }
    Poly permutation_product(const std::vector<const Poly*>& values, const std::vector<const Poly*>& sigma, const Fr& beta, const Fr& gamma,
                             const Fr& delta_omega_start, const Fr& delta, const Fr& z_init) override {
        std::vector<const fr_t*> tv, ts;
        for (auto* c : values) tv.push_back(fr(*c));
        for (auto* c : sigma) ts.push_back(fr(*c));
        Poly z((size_t)dom_.n);
        if (halo2_permutation_product(tv.data(), ts.data(), (uint32_t)tv.size(), fr1(beta), fr1(gamma), fr1(delta_omega_start), fr1(delta),
                                      &dom_.omega, dom_.k, fr1(z_init), fr(z)) != 0)
            throw Panic("halo2_permutation_product failed");
        return z;
    }
    Poly logup_running_sum(const std::vector<const Poly*>& inputs, const Poly& table, const Poly& m, const Fr& beta, const Fr& phi_init) override {
        std::vector<const fr_t*> ti;
        for (auto* c : inputs) ti.push_back(fr(*c));
        Poly phi((size_t)dom_.n);
        if (halo2_logup_running_sum(ti.data(), (uint32_t)ti.size(), fr(table), fr(m), fr1(beta), dom_.k, fr1(phi_init), fr(phi)) != 0)
            throw Panic("halo2_logup_running_sum failed");
        return phi;
    }

  private:
    const std::vector<G1Affine>& g_;
    const std::vector<G1Affine>& gl_;
    halo2_domain_t dom_;
};
}  // namespace oracle_ops
