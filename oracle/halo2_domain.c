/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Restates halo2_proofs 1.1.0 (scroll-tech/halo2 @ e5ddf67, pin
 * /root/reference/Cargo.lock:1886-1888):
 *   halo2_proofs/src/poly/domain.rs   EvaluationDomain::{new, lagrange_to_coeff,
 *       coeff_to_extended, extended_to_coeff, distribute_powers_zeta, ifft}
 *   halo2_proofs/src/poly/kzg/commitment.rs   ParamsKZG::{setup, commit, commit_lagrange},
 *       g_to_lagrange (used by Params::downsize, reference call site
 *       /root/reference/integration/tests/integration.rs:17-18)
 * The k = 25 domain constants produced by halo2_domain_new are pinned against
 * /root/reference/release-v0.13.1/chunk.protocol by tests/test_oracle_fixtures.py.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "bn254_oracle.h"

int halo2_domain_new(halo2_domain_t *d, uint32_t j, uint32_t k) {
    memset(d, 0, sizeof *d);
    if (j < 2 || k > FR_S) return -1;
    uint64_t qpd = (uint64_t)j - 1;
    uint64_t n = 1ULL << k;
    uint32_t ek = k;
    while ((1ULL << ek) < n * qpd) ek++;
    if (ek > FR_S || (ek - k) > 6) return -1;
    d->n = n;
    d->k = k;
    d->extended_k = ek;
    d->quotient_poly_degree = (uint32_t)qpd;

    fr_t eo = fr_ROOT_OF_UNITY;
    for (uint32_t i = ek; i < FR_S; ++i) fr_sqr(&eo, &eo);
    fr_t om = eo;
    for (uint32_t i = k; i < ek; ++i) fr_sqr(&om, &om);
    d->extended_omega = eo;
    d->omega = om;
    fr_inv(&d->extended_omega_inv, &eo);
    fr_inv(&d->omega_inv, &om);
    d->g_coset = fr_ZETA;
    fr_sqr(&d->g_coset_inv, &fr_ZETA);

    /* t(X) = X^n - 1 on the coset zeta * <extended_omega>; repeats with period 2^(ek-k) */
    uint64_t e[4] = {n, 0, 0, 0};
    fr_t orig, step, cur;
    fr_pow(&orig, &fr_ZETA, e);
    fr_pow(&step, &eo, e);
    cur = orig;
    uint32_t cnt = 0;
    do {
        d->t_evaluations[cnt++] = cur;
        fr_mul(&cur, &cur, &step);
    } while (!fr_eq(&cur, &orig) && cnt < 64);
    if (cnt != (1u << (ek - k))) return -2;
    d->n_t_evaluations = cnt;
    for (uint32_t i = 0; i < cnt; ++i) {
        fr_sub(&d->t_evaluations[i], &d->t_evaluations[i], &fr_ONE);
        fr_inv(&d->t_evaluations[i], &d->t_evaluations[i]);
    }
    fr_t t;
    fr_from_u64(&t, 1ULL << k);
    fr_inv(&d->ifft_divisor, &t);
    fr_from_u64(&t, 1ULL << ek);
    fr_inv(&d->extended_ifft_divisor, &t);
    fr_from_u64(&t, n);
    fr_inv(&d->barycentric_weight, &t);
    return 0;
}

/* a[i] *= {1, g, g^2}[i % 3] with g = zeta (into coset) or zeta^-1 = zeta^2 (out of coset) */
void halo2_distribute_powers_zeta(const halo2_domain_t *d, fr_t *a, uint64_t len, int into_coset) {
    const fr_t *p1 = into_coset ? &d->g_coset : &d->g_coset_inv;
    const fr_t *p2 = into_coset ? &d->g_coset_inv : &d->g_coset;
    for (uint64_t i = 0; i < len; ++i) {
        uint64_t m = i % 3;
        if (m == 1)
            fr_mul(&a[i], &a[i], p1);
        else if (m == 2)
            fr_mul(&a[i], &a[i], p2);
    }
}

/* EvaluationDomain::ifft */
static void domain_ifft(fr_t *a, const fr_t *omega_inv, uint32_t log_n, const fr_t *divisor, int threads) {
    halo2_best_fft(a, omega_inv, log_n, threads);
    uint64_t n = 1ULL << log_n;
    for (uint64_t i = 0; i < n; ++i) fr_mul(&a[i], &a[i], divisor);
}

void halo2_lagrange_to_coeff(const halo2_domain_t *d, fr_t *a, int threads) {
    domain_ifft(a, &d->omega_inv, d->k, &d->ifft_divisor, threads);
}

void halo2_coeff_to_extended(const halo2_domain_t *d, const fr_t *a, fr_t *out, int threads) {
    uint64_t en = 1ULL << d->extended_k;
    memcpy(out, a, sizeof(fr_t) * d->n);
    halo2_distribute_powers_zeta(d, out, d->n, 1);
    memset(out + d->n, 0, sizeof(fr_t) * (en - d->n));
    halo2_best_fft(out, &d->extended_omega, d->extended_k, threads);
}

void halo2_extended_to_coeff(const halo2_domain_t *d, fr_t *a, int threads) {
    uint64_t en = 1ULL << d->extended_k;
    domain_ifft(a, &d->extended_omega_inv, d->extended_k, &d->extended_ifft_divisor, threads);
    halo2_distribute_powers_zeta(d, a, en, 0);
    /* caller truncates to n * quotient_poly_degree */
}

/* ------------------------------------------------------------------ ParamsKZG */

typedef struct {
    const fr_t *scalars;
    g1_affine_t *out;
    uint64_t lo, hi;
} smul_job_t;

static void *smul_worker(void *arg) {
    smul_job_t *j = (smul_job_t *)arg;
    g1_affine_t gen;
    g1_generator(&gen);
    g1_t G;
    g1_from_affine(&G, &gen);
    uint64_t cnt = j->hi - j->lo;
    g1_t *tmp = (g1_t *)malloc(sizeof(g1_t) * (cnt ? cnt : 1));
    for (uint64_t i = j->lo; i < j->hi; ++i) g1_mul(&tmp[i - j->lo], &G, &j->scalars[i]);
    g1_batch_normalize(j->out + j->lo, tmp, cnt);
    free(tmp);
    return NULL;
}

static void generator_mul_batch(const fr_t *scalars, g1_affine_t *out, uint64_t n, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    smul_job_t *jobs = (smul_job_t *)malloc(sizeof(smul_job_t) * threads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (smul_job_t){scalars, out, n * t / threads, n * (t + 1) / threads};
        pthread_create(&tids[t], NULL, smul_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
}

/* ParamsKZG::setup with a caller-chosen tau ("unsafe" SRS): g[i] = tau^i G,
 * g_lagrange[i] = L_i(tau) G, L_i(tau) = (tau^n - 1)/n * w^i / (tau - w^i) */
void halo2_params_setup(uint32_t k, const fr_t *tau, g1_affine_t *g, g1_affine_t *g_lagrange, int threads) {
    uint64_t n = 1ULL << k;
    fr_t *sc = (fr_t *)malloc(sizeof(fr_t) * n);
    fr_t cur = fr_ONE;
    for (uint64_t i = 0; i < n; ++i) {
        sc[i] = cur;
        fr_mul(&cur, &cur, tau);
    }
    generator_mul_batch(sc, g, n, threads);
    if (g_lagrange) {
        fr_t root = fr_ROOT_OF_UNITY, n_inv, nf, mult, rp = fr_ONE;
        for (uint32_t i = k; i < FR_S; ++i) fr_sqr(&root, &root);
        fr_from_u64(&nf, n);
        fr_inv(&n_inv, &nf);
        /* cur == tau^n here */
        fr_sub(&mult, &cur, &fr_ONE);
        fr_mul(&mult, &mult, &n_inv);
        for (uint64_t i = 0; i < n; ++i) {
            fr_t den;
            fr_sub(&den, tau, &rp);
            fr_inv(&den, &den);
            fr_mul(&sc[i], &mult, &rp);
            fr_mul(&sc[i], &sc[i], &den);
            fr_mul(&rp, &rp, &root);
        }
        generator_mul_batch(sc, g_lagrange, n, threads);
    }
    free(sc);
}

/* g_to_lagrange: inverse FFT over G1 then * n^-1 then batch_normalize */
void halo2_g_to_lagrange(const g1_affine_t *g, g1_affine_t *g_lagrange, uint32_t k, int threads) {
    uint64_t n = 1ULL << k;
    g1_t *p = (g1_t *)malloc(sizeof(g1_t) * n);
    for (uint64_t i = 0; i < n; ++i) g1_from_affine(&p[i], &g[i]);
    fr_t root = fr_ROOT_OF_UNITY, omega_inv, nf, n_inv;
    for (uint32_t i = k; i < FR_S; ++i) fr_sqr(&root, &root);
    fr_inv(&omega_inv, &root);
    fr_from_u64(&nf, n);
    fr_inv(&n_inv, &nf);
    halo2_best_fft_g1(p, &omega_inv, k, threads);
    for (uint64_t i = 0; i < n; ++i) g1_mul(&p[i], &p[i], &n_inv);
    g1_batch_normalize(g_lagrange, p, n);
    free(p);
}

/* ParamsKZG::commit / commit_lagrange: best_multiexp over the first n bases; Blind ignored */
void halo2_commit(const g1_affine_t *bases, const fr_t *poly, uint64_t n, int threads, g1_t *out) {
    halo2_best_multiexp(poly, bases, n, threads, out);
}
