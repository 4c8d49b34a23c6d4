// GPU test of the C++ host mirror (scroll-prover_b200/halo2_b200.hpp) through the C ABI.
// Mirrors the style of halo2_proofs' own unit tests: commit(p) == [p(s)] G on a test SRS, fft round trips,
// coset extension round trip, params file round trip, panics on length mismatch.
// Expected values come from the host emulation of ff.cuh / ec.cuh (double-and-add, Horner) -- a different
// algorithm than the device Pippenger / NTT.
#include <cstdio>
#include <cstdlib>

#include "../../scroll-prover_b200/halo2_b200.hpp"
#include "../../scroll-prover_b200/csrc/ec.cuh"

using namespace halo2_b200;
using detail::DFr;

#define REQUIRE(c)                                                     \
    do {                                                               \
        if (!(c)) {                                                    \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            return 1;                                                  \
        }                                                              \
    } while (0)

static uint64_t rng_state = 0x5EEDB2000001ull;
static uint64_t xs() {
    uint64_t x = rng_state;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    rng_state = x;
    return x * 0x2545F4914F6CDD1Dull;
}
static Fr rand_fr() {
    Fr r{{xs(), xs(), xs(), xs() & 0x0fffffffffffffffull}};  // < 2^252 < r: valid Montgomery limbs
    return r;
}
// [s] G by host double-and-add (s in Montgomery form), affine
static G1Affine host_generator_mul(const Fr& s) {
    b200zk::Fr c = detail::to_dev(s).from_mont();
    b200zk::Fq gx = b200zk::Fq::one(), gy = b200zk::Fq::one().dbl();
    b200zk::XYZZ acc = b200zk::XYZZ::identity();
    for (int limb = 7; limb >= 0; --limb)
        for (int b = 31; b >= 0; --b) {
            acc = b200zk::xyzz_dbl(acc);
            if ((c.l.v[limb] >> b) & 1) b200zk::xyzz_madd(acc, gx, gy);
        }
    b200zk::Affine a = b200zk::xyzz_to_affine(acc);
    G1Affine out;
    std::memcpy(&out, &a, 64);
    return out;
}
static Fr host_eval(const std::vector<Fr>& p, const Fr& x) {
    DFr acc = DFr::zero(), dx = detail::to_dev(x);
    for (size_t i = p.size(); i-- > 0;) acc = acc * dx + detail::to_dev(p[i]);
    return detail::from_dev(acc);
}
static bool same_point(const G1& j, const G1Affine& a) {  // device results are normalised (x, y, 1)
    return std::memcmp(&j.x, &a.x, 32) == 0 && std::memcmp(&j.y, &a.y, 32) == 0 && !j.is_identity();
}

int main() {
    const uint32_t k = 10;
    const size_t n = size_t(1) << k;
    EvaluationDomain dom = EvaluationDomain::new_(5, k);
    REQUIRE(dom.extended_k == k + 2 && dom.quotient_poly_degree == 4);
    REQUIRE(detail::from_dev(detail::to_dev(dom.omega).pow_u64(n)) == detail::from_dev(DFr::one()));
    REQUIRE(detail::from_dev(detail::to_dev(dom.omega) * detail::to_dev(dom.omega_inv)) == detail::from_dev(DFr::one()));

    ParamsKZG params;
    Fr s = rand_fr();
    ParamsKZG::setup(params, k, s);
    REQUIRE(params.g.size() == n && params.g_lagrange.size() == n);
    {
        G1Affine g1 = host_generator_mul(s);  // g[1] = [s] G
        REQUIRE(std::memcmp(&params.g[1], &g1, 64) == 0);
    }
    std::vector<Fr> poly(n);
    for (auto& c : poly) c = rand_fr();
    poly[0] = Fr{{0, 0, 0, 0}};

    // commit(p) == [p(s)] G
    G1 c1 = params.commit(poly);
    Fr ps = host_eval(poly, s);
    REQUIRE(arithmetic::eval_polynomial(poly, s) == ps);
    REQUIRE(same_point(c1, host_generator_mul(ps)));

    // commit_lagrange(evals) == commit(coeffs), evals = best_fft(coeffs, omega)
    std::vector<Fr> evals = poly;
    arithmetic::best_fft(evals, dom.omega, k);
    G1 c2 = params.commit_lagrange(evals);
    REQUIRE(std::memcmp(&c1, &c2, 96) == 0);
    // A[1] = p(omega)
    REQUIRE(evals[1] == host_eval(poly, dom.omega));

    // lagrange_to_coeff inverts it
    std::vector<Fr> back = dom.lagrange_to_coeff(evals);
    REQUIRE(back == poly);

    // coeff_to_extended: ext[i] = p(zeta * w_ext^i); extended_to_coeff returns the (zero-padded) coefficients
    std::vector<Fr> ext = dom.coeff_to_extended(poly);
    REQUIRE(ext.size() == 4 * n);
    Fr pt = detail::from_dev(detail::to_dev(dom.g_coset) * detail::to_dev(dom.extended_omega).pow_u64(5));
    REQUIRE(ext[5] == host_eval(poly, pt));
    std::vector<Fr> coeffs4 = dom.extended_to_coeff(ext);
    REQUIRE(coeffs4.size() == 4 * n);
    for (size_t i = 0; i < n; ++i) REQUIRE(coeffs4[i] == poly[i]);
    for (size_t i = n; i < 4 * n; ++i) REQUIRE((coeffs4[i].l[0] | coeffs4[i].l[1] | coeffs4[i].l[2] | coeffs4[i].l[3]) == 0);

    // batched column pipeline: commitments equal commit_lagrange; device coefficient form equals lagrange_to_coeff
    {
        std::vector<Fr> col2(n);
        for (auto& c : col2) c = rand_fr();
        void* coeff_dev[2] = {nullptr, nullptr};
        REQUIRE(b200zk_buf_alloc(Backend::get().ctx(), 32 * n, &coeff_dev[0]) == B200ZK_OK);
        REQUIRE(b200zk_buf_alloc(Backend::get().ctx(), 32 * n, &coeff_dev[1]) == B200ZK_OK);
        std::vector<const Fr*> cols = {evals.data(), col2.data()};
        std::vector<G1> cm = commit_columns(params, dom, cols, 1, coeff_dev, nullptr);
        G1 e0 = params.commit_lagrange(evals), e1 = params.commit_lagrange(col2);
        REQUIRE(std::memcmp(&cm[0], &e0, 96) == 0 && std::memcmp(&cm[1], &e1, 96) == 0);
        std::vector<Fr> got(n);
        REQUIRE(b200zk_buf_download(Backend::get().ctx(), got.data(), coeff_dev[0], 32 * n) == B200ZK_OK);
        REQUIRE(got == poly);  // evals = NTT(poly)
        b200zk_buf_free(Backend::get().ctx(), coeff_dev[0]);
        b200zk_buf_free(Backend::get().ctx(), coeff_dev[1]);
    }

    // best_multiexp: generic bases, shorter slices, and the reference's length assertion
    std::vector<Fr> sc(poly.begin(), poly.begin() + 100);
    std::vector<G1Affine> bs(params.g.begin(), params.g.begin() + 100);
    G1 m1 = arithmetic::best_multiexp(sc, bs);
    G1 m2 = params.commit(sc);
    REQUIRE(std::memcmp(&m1, &m2, 96) == 0);
    bool panicked = false;
    try {
        bs.pop_back();
        arithmetic::best_multiexp(sc, bs);
    } catch (const Panic&) {
        panicked = true;
    }
    REQUIRE(panicked);
    panicked = false;
    try {
        std::vector<Fr> wrong(n / 2);
        arithmetic::best_fft(wrong, dom.omega, k);
    } catch (const Panic&) {
        panicked = true;
    }
    REQUIRE(panicked);

    // kate_division: a(X) - a(b) = q(X) (X - b)  => check at a random point
    Fr bpt = rand_fr(), x = rand_fr();
    std::vector<Fr> q = arithmetic::kate_division(poly, bpt);
    DFr lhs = detail::to_dev(host_eval(poly, x)) - detail::to_dev(host_eval(poly, bpt));
    DFr rhs = detail::to_dev(host_eval(q, x)) * (detail::to_dev(x) - detail::to_dev(bpt));
    REQUIRE(detail::from_dev(lhs) == detail::from_dev(rhs));

    // params file round trip (SerdeFormat::RawBytes layout)
    const char* path = "/tmp/b200zk_params_test.bin";
    params.write_custom(path);
    ParamsKZG p2;
    ParamsKZG::read_custom(p2, path);
    REQUIRE(p2.k == k && p2.g.size() == n && std::memcmp(p2.g.data(), params.g.data(), 64 * n) == 0 &&
            std::memcmp(p2.g_lagrange.data(), params.g_lagrange.data(), 64 * n) == 0);
    G1 c3 = p2.commit(poly);
    REQUIRE(std::memcmp(&c1, &c3, 96) == 0);
    std::remove(path);

    // Params::downsize: g_lagrange of the smaller domain is rebuilt by the device G1 FFT; commit == commit_lagrange again
    p2.downsize(k - 2);
    {
        EvaluationDomain d2 = EvaluationDomain::new_(5, k - 2);
        std::vector<Fr> small(poly.begin(), poly.begin() + (n >> 2));
        G1 a1 = p2.commit(small);
        std::vector<Fr> ev = small;
        arithmetic::best_fft(ev, d2.omega, k - 2);
        G1 a2 = p2.commit_lagrange(ev);
        REQUIRE(std::memcmp(&a1, &a2, 96) == 0);
    }

    // plonk::GraphEvaluator on device-resident extended columns: h = h*y + q*(a(X)*a(wX) - b), rows checked on the host
    {
        using plonk::ValueSource;
        const size_t en = 4 * n;
        std::vector<Fr> a(en), bcol(en), qf(en), h(en);
        for (size_t i = 0; i < en; ++i) { a[i] = rand_fr(); bcol[i] = rand_fr(); qf[i] = rand_fr(); h[i] = rand_fr(); }
        DeviceColumn da(a), db(bcol), dq(qf), dh(h);
        plonk::GraphEvaluator ev;
        uint32_t r0 = ev.add_rotation(0), r1 = ev.add_rotation(1);
        ValueSource m = ev.add(B200ZK_CALC_MUL, ValueSource::Advice(0, r0), ValueSource::Advice(0, r1));
        ValueSource d = ev.add(B200ZK_CALC_SUB, m, ValueSource::Advice(1, r0));
        ValueSource g = ev.add(B200ZK_CALC_MUL, d, ValueSource::Fixed(0, r0));
        ev.add_horner(ValueSource::PreviousValue(), {g}, ValueSource::Y());
        Fr y = rand_fr(), zero{{0, 0, 0, 0}};
        ev.evaluate(dh, dom, {&dq}, {&da, &db}, {}, {}, zero, zero, zero, y);
        std::vector<Fr> got = dh.to_host();
        for (size_t i : {size_t(0), size_t(1), en / 2 + 3, en - 4, en - 1}) {
            size_t nxt = (i + 4) % en;  // rotation 1 of the original domain = 4 rows of the extended one
            DFr want = detail::to_dev(h[i]) * detail::to_dev(y) +
                       (detail::to_dev(a[i]) * detail::to_dev(a[nxt]) - detail::to_dev(bcol[i])) * detail::to_dev(qf[i]);
            REQUIRE(got[i] == detail::from_dev(want));
        }
        // permutation z for the identity permutation (sigma = labels): every ratio is 1, so z stays at z_init
        std::vector<Fr> v(n), lab(n);
        DFr wpow = DFr::one();
        for (size_t i = 0; i < n; ++i) {
            v[i] = rand_fr();
            lab[i] = detail::from_dev(wpow);
            wpow = wpow * detail::to_dev(dom.omega);
        }
        DeviceColumn dv(v), dl(lab), dz(n);
        Fr one = detail::from_dev(DFr::one()), seven = detail::from_dev(detail::from_u64(7)), z0 = rand_fr();
        plonk::permutation_product({&dv}, {&dl}, rand_fr(), rand_fr(), one, seven, dom, z0, dz);
        std::vector<Fr> z = dz.to_host();
        REQUIRE(z[0] == z0 && z[n / 2] == z0 && z[n - 1] == z0);

        // evaluate_h's permutation and lookup sections as generated programs, checked on a few rows against the
        // upstream formulas evaluated with the host field emulation
        auto rnd_col = [&]() { std::vector<Fr> c(en); for (auto& x : c) x = rand_fr(); return c; };
        std::vector<Fr> zc = rnd_col(), v0 = rnd_col(), v1 = rnd_col(), s0 = rnd_col(), s1 = rnd_col(), l0c = rnd_col(),
                        llc = rnd_col(), lac = rnd_col(), prev = rnd_col();
        DeviceColumn dzc(zc), dv0(v0), dv1(v1), ds0(s0), ds1(s1), dl0(l0c), dll(llc), dla(lac), dprev(prev);
        Fr beta = rand_fr(), gamma = rand_fr(), delta = detail::from_dev(detail::from_u64(7).pow_u64(1ull << 28));
        const int32_t last_rot = -6;
        {
            plonk::GraphEvaluator pe;
            uint32_t q0 = pe.add_rotation(0);
            plonk::permutation_constraints(pe, {ValueSource::Advice(0, q0)}, 3, {ValueSource::Advice(1, q0), ValueSource::Advice(2, q0)},
                                           {ValueSource::Fixed(0, q0), ValueSource::Fixed(1, q0)}, ValueSource::Fixed(2, q0),
                                           ValueSource::Fixed(3, q0), ValueSource::Fixed(4, q0), last_rot, delta);
            pe.evaluate(dprev, dom, {&ds0, &ds1, &dl0, &dll, &dla}, {&dzc, &dv0, &dv1}, {}, {}, beta, gamma, zero, y);
            std::vector<Fr> outp = dprev.to_host();
            auto D = [](const Fr& f) { return detail::to_dev(f); };
            for (size_t i : {size_t(0), size_t(7), en - 1}) {
                size_t nxt = (i + 4) % en;
                DFr X = D(dom.g_coset) * D(dom.extended_omega).pow_u64(i), dy = D(y), b_ = D(beta), g_ = D(gamma);
                DFr val = D(prev[i]);
                val = val * dy + (DFr::one() - D(zc[i])) * D(l0c[i]);
                val = val * dy + (D(zc[i]) * D(zc[i]) - D(zc[i])) * D(llc[i]);
                DFr left = D(zc[nxt]) * (D(v0[i]) + b_ * D(s0[i]) + g_) * (D(v1[i]) + b_ * D(s1[i]) + g_);
                DFr right = D(zc[i]) * (D(v0[i]) + b_ * X + g_) * (D(v1[i]) + b_ * X * D(delta) + g_);
                val = val * dy + (left - right) * D(lac[i]);
                REQUIRE(outp[i] == detail::from_dev(val));
            }
        }
        {
            std::vector<Fr> f0 = rnd_col(), f1 = rnd_col(), tb = rnd_col(), mm = rnd_col(), ph = rnd_col(), prev2 = rnd_col();
            DeviceColumn df0(f0), df1(f1), dtb(tb), dmm(mm), dph(ph), dprev2(prev2);
            plonk::GraphEvaluator le;
            uint32_t q0 = le.add_rotation(0);
            plonk::lookup_constraints(le, {ValueSource::Advice(0, q0), ValueSource::Advice(1, q0)}, ValueSource::Advice(2, q0),
                                      ValueSource::Advice(3, q0), ValueSource::Advice(4, q0), ValueSource::Fixed(0, q0),
                                      ValueSource::Fixed(1, q0), ValueSource::Fixed(2, q0));
            le.evaluate(dprev2, dom, {&dl0, &dll, &dla}, {&df0, &df1, &dtb, &dmm, &dph}, {}, {}, beta, zero, zero, y);
            std::vector<Fr> outl = dprev2.to_host();
            auto D = [](const Fr& f) { return detail::to_dev(f); };
            for (size_t i : {size_t(0), size_t(9), en - 2}) {
                size_t nxt = (i + 4) % en;
                DFr dy = D(y), b_ = D(beta);
                DFr p0 = D(f0[i]) + b_, p1 = D(f1[i]) + b_, tau = D(tb[i]) + b_, prod = p0 * p1;
                DFr lhs = tau * prod * (D(ph[nxt]) - D(ph[i]));
                DFr rhs = prod * (tau * (p0.inv() + p1.inv()) - D(mm[i]));  // upstream's form, with inversions
                DFr val = D(prev2[i]);
                val = val * dy + D(l0c[i]) * D(ph[i]);
                val = val * dy + D(llc[i]) * D(ph[i]);
                val = val * dy + (lhs - rhs) * D(lac[i]);
                REQUIRE(outl[i] == detail::from_dev(val));
            }
        }
    }

    std::printf("ALL OK\n");
    return 0;
}
