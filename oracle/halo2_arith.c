/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * Restates halo2_proofs 1.1.0 (scroll-tech/halo2 @ e5ddf67, pin
 * /root/reference/Cargo.lock:1886-1888) halo2_proofs/src/arithmetic.rs:
 *   multiexp_serial, best_multiexp, best_fft, recursive_butterfly_arithmetic,
 *   eval_polynomial, kate_division, compute_inner_product.
 * Reached from the reference at integration/src/prove.rs:37-39,67,95-97.
 * Rayon's multicore::scope / join is mirrored with pthreads (same work split).
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "bn254_oracle.h"

/* ------------------------------------------------------------------ MSM */

/* get_at(segment, c, bytes): c-bit window #segment of the 32-byte LE repr */
static inline uint64_t get_at(uint32_t segment, uint32_t c, const uint8_t *bytes) {
    uint32_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
    if (skip_bytes >= 32) return 0;
    uint8_t v[8] = {0};
    uint32_t avail = 32 - skip_bytes;
    memcpy(v, bytes + skip_bytes, avail < 8 ? avail : 8);
    uint64_t tmp;
    memcpy(&tmp, v, 8);
    tmp >>= skip_bits - skip_bytes * 8;
    return tmp % (1ULL << c);
}

typedef struct {
    int kind; /* 0 None, 1 Affine, 2 Projective */
    g1_affine_t a;
    g1_t p;
} bucket_t;

static uint32_t window_for(uint64_t n) {
    if (n < 4) return 1;
    if (n < 32) return 3;
    return (uint32_t)ceil(log((double)(uint32_t)n));
}

void halo2_multiexp_serial(const fr_t *coeffs, const g1_affine_t *bases, uint64_t n, g1_t *acc) {
    uint8_t *repr = (uint8_t *)malloc(32 * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) fr_to_repr(repr + 32 * i, &coeffs[i]);
    uint32_t c = window_for(n);
    uint32_t segments = 256 / c + 1;
    uint64_t nb = (1ULL << c) - 1;
    bucket_t *buckets = (bucket_t *)malloc(sizeof(bucket_t) * nb);
    for (uint32_t seg = segments; seg-- > 0;) {
        for (uint32_t i = 0; i < c; ++i) g1_double(acc, acc);
        for (uint64_t b = 0; b < nb; ++b) buckets[b].kind = 0;
        for (uint64_t i = 0; i < n; ++i) {
            uint64_t d = get_at(seg, c, repr + 32 * i);
            if (!d) continue;
            bucket_t *bk = &buckets[d - 1];
            if (bk->kind == 0) {
                bk->kind = 1;
                bk->a = bases[i];
            } else if (bk->kind == 1) {
                g1_t t;
                g1_from_affine(&t, &bk->a);
                g1_add_mixed(&bk->p, &t, &bases[i]);
                bk->kind = 2;
            } else {
                g1_add_mixed(&bk->p, &bk->p, &bases[i]);
            }
        }
        /* summation by parts */
        g1_t running;
        g1_identity(&running);
        for (uint64_t b = nb; b-- > 0;) {
            if (buckets[b].kind == 1)
                g1_add_mixed(&running, &running, &buckets[b].a);
            else if (buckets[b].kind == 2)
                g1_add(&running, &running, &buckets[b].p);
            g1_add(acc, acc, &running);
        }
    }
    free(buckets);
    free(repr);
}

typedef struct {
    const fr_t *coeffs;
    const g1_affine_t *bases;
    uint64_t n;
    g1_t acc;
} msm_job_t;

static void *msm_worker(void *arg) {
    msm_job_t *j = (msm_job_t *)arg;
    g1_identity(&j->acc);
    halo2_multiexp_serial(j->coeffs, j->bases, j->n, &j->acc);
    return NULL;
}

void halo2_best_multiexp(const fr_t *coeffs, const g1_affine_t *bases, uint64_t n, int threads, g1_t *out) {
    if (threads < 1) threads = 1;
    if (n > (uint64_t)threads) {
        uint64_t chunk = n / (uint64_t)threads;
        uint64_t num_chunks = (n + chunk - 1) / chunk;
        msm_job_t *jobs = (msm_job_t *)malloc(sizeof(msm_job_t) * num_chunks);
        pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * num_chunks);
        for (uint64_t i = 0; i < num_chunks; ++i) {
            uint64_t off = i * chunk;
            jobs[i].coeffs = coeffs + off;
            jobs[i].bases = bases + off;
            jobs[i].n = (off + chunk <= n) ? chunk : n - off;
            if (num_chunks == 1)
                msm_worker(&jobs[i]);
            else
                pthread_create(&tids[i], NULL, msm_worker, &jobs[i]);
        }
        g1_identity(out);
        for (uint64_t i = 0; i < num_chunks; ++i) {
            if (num_chunks > 1) pthread_join(tids[i], NULL);
            g1_add(out, out, &jobs[i].acc);
        }
        free(jobs);
        free(tids);
    } else {
        g1_identity(out);
        halo2_multiexp_serial(coeffs, bases, n, out);
    }
}

/* ------------------------------------------------------------------ FFT (FftGroup = Fr or G1) */

static uint64_t bitreverse(uint64_t n, uint32_t l) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < l; ++i) {
        r = (r << 1) | (n & 1);
        n >>= 1;
    }
    return r;
}

static uint32_t log2_floor(uint32_t num) {
    uint32_t pow = 0;
    while ((1u << (pow + 1)) <= num) pow++;
    return pow;
}

#define FFT_IMPL(SUF, G, G_ADD, G_SUB, G_SCALE)                                                          \
    typedef struct {                                                                                     \
        G *a;                                                                                            \
        uint64_t n, twiddle_chunk;                                                                       \
        const fr_t *twiddles;                                                                            \
        int depth;                                                                                       \
    } rec_job_##SUF;                                                                                     \
    static void recursive_butterfly_##SUF(G *a, uint64_t n, uint64_t twiddle_chunk,                      \
                                          const fr_t *twiddles, int depth);                              \
    static void *rec_worker_##SUF(void *arg) {                                                           \
        rec_job_##SUF *j = (rec_job_##SUF *)arg;                                                         \
        recursive_butterfly_##SUF(j->a, j->n, j->twiddle_chunk, j->twiddles, j->depth);                  \
        return NULL;                                                                                     \
    }                                                                                                    \
    static inline void butterfly_pair_##SUF(G *a, G *b, const fr_t *w) {                                 \
        G t = *b;                                                                                        \
        if (w) G_SCALE(&t, &t, w);                                                                       \
        *b = *a;                                                                                         \
        G_ADD(a, a, &t);                                                                                 \
        G_SUB(b, b, &t);                                                                                 \
    }                                                                                                    \
    static void recursive_butterfly_##SUF(G *a, uint64_t n, uint64_t twiddle_chunk,                      \
                                          const fr_t *twiddles, int depth) {                             \
        if (n == 2) {                                                                                    \
            butterfly_pair_##SUF(&a[0], &a[1], NULL);                                                    \
            return;                                                                                      \
        }                                                                                                \
        G *left = a, *right = a + n / 2;                                                                 \
        if (depth > 0) { /* multicore::join */                                                           \
            rec_job_##SUF j = {left, n / 2, twiddle_chunk * 2, twiddles, depth - 1};                     \
            pthread_t tid;                                                                               \
            pthread_create(&tid, NULL, rec_worker_##SUF, &j);                                            \
            recursive_butterfly_##SUF(right, n / 2, twiddle_chunk * 2, twiddles, depth - 1);             \
            pthread_join(tid, NULL);                                                                     \
        } else {                                                                                         \
            recursive_butterfly_##SUF(left, n / 2, twiddle_chunk * 2, twiddles, 0);                      \
            recursive_butterfly_##SUF(right, n / 2, twiddle_chunk * 2, twiddles, 0);                     \
        }                                                                                                \
        butterfly_pair_##SUF(&left[0], &right[0], NULL); /* twiddle factor one */                        \
        for (uint64_t i = 1; i < n / 2; ++i)                                                             \
            butterfly_pair_##SUF(&left[i], &right[i], &twiddles[i * twiddle_chunk]);                     \
    }                                                                                                    \
    void halo2_best_fft##SUF(G *a, const fr_t *omega, uint32_t log_n, int threads) {                     \
        if (threads < 1) threads = 1;                                                                    \
        uint32_t log_threads = log2_floor((uint32_t)threads);                                            \
        uint64_t n = 1ULL << log_n;                                                                      \
        for (uint64_t k = 0; k < n; ++k) {                                                               \
            uint64_t rk = bitreverse(k, log_n);                                                          \
            if (k < rk) {                                                                                \
                G t = a[rk];                                                                             \
                a[rk] = a[k];                                                                            \
                a[k] = t;                                                                                \
            }                                                                                            \
        }                                                                                                \
        uint64_t nt = n / 2 ? n / 2 : 1;                                                                 \
        fr_t *twiddles = (fr_t *)malloc(sizeof(fr_t) * nt);                                              \
        fr_t w = fr_ONE;                                                                                 \
        for (uint64_t i = 0; i < n / 2; ++i) {                                                           \
            twiddles[i] = w;                                                                             \
            fr_mul(&w, &w, omega);                                                                       \
        }                                                                                                \
        if (log_n <= log_threads) {                                                                      \
            uint64_t chunk = 2, twiddle_chunk = n / 2;                                                   \
            for (uint32_t s = 0; s < log_n; ++s) {                                                       \
                for (uint64_t base = 0; base < n; base += chunk) {                                       \
                    G *left = a + base, *right = a + base + chunk / 2;                                   \
                    butterfly_pair_##SUF(&left[0], &right[0], NULL);                                     \
                    for (uint64_t i = 1; i < chunk / 2; ++i)                                             \
                        butterfly_pair_##SUF(&left[i], &right[i], &twiddles[i * twiddle_chunk]);         \
                }                                                                                        \
                chunk *= 2;                                                                              \
                twiddle_chunk /= 2;                                                                      \
            }                                                                                            \
        } else if (n >= 2) {                                                                             \
            recursive_butterfly_##SUF(a, n, 1, twiddles, (int)log_threads);                              \
        }                                                                                                \
        free(twiddles);                                                                                  \
    }

static inline void g1_sub_(g1_t *r, const g1_t *a, const g1_t *b) {
    g1_t nb;
    g1_neg(&nb, b);
    g1_add(r, a, &nb);
}
static inline void g1_scale_(g1_t *r, const g1_t *a, const fr_t *s) { g1_mul(r, a, s); }

FFT_IMPL(, fr_t, fr_add, fr_sub, fr_mul)
FFT_IMPL(_g1, g1_t, g1_add, g1_sub_, g1_scale_)

/* ------------------------------------------------------------------ poly helpers */

void halo2_eval_polynomial(fr_t *r, const fr_t *poly, uint64_t n, const fr_t *point) {
    fr_t acc;
    memset(&acc, 0, sizeof acc);
    for (uint64_t i = n; i-- > 0;) {
        fr_mul(&acc, &acc, point);
        fr_add(&acc, &acc, &poly[i]);
    }
    *r = acc;
}

/* divide a(X) (n coeffs) by (X - b); q gets n-1 coeffs, remainder dropped */
void halo2_kate_division(fr_t *q, const fr_t *a, uint64_t n, const fr_t *b) {
    fr_t nb, tmp;
    fr_neg(&nb, b);
    memset(&tmp, 0, sizeof tmp);
    if (n < 2) return;
    for (uint64_t i = n - 1; i >= 1; --i) {
        fr_t lead;
        fr_sub(&lead, &a[i], &tmp);
        q[i - 1] = lead;
        fr_mul(&tmp, &lead, &nb);
    }
}

void halo2_compute_inner_product(fr_t *r, const fr_t *a, const fr_t *b, uint64_t n) {
    fr_t acc, t;
    memset(&acc, 0, sizeof acc);
    for (uint64_t i = 0; i < n; ++i) {
        fr_mul(&t, &a[i], &b[i]);
        fr_add(&acc, &acc, &t);
    }
    *r = acc;
}

/* ------------------------------------------------------------------ deterministic test vectors */

static inline uint64_t xs64star(uint64_t *s) {
    uint64_t x = *s;
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    *s = x;
    return x * 0x2545F4914F6CDD1DULL;
}

/* xorshift64* stream, 4 limbs, top masked to 254 bits, rejection-sampled < r,
 * stored Montgomery. witness_like: 60 % zero, 30 % < 2^16, 10 % uniform. */
void oracle_fill_fr(fr_t *out, uint64_t n, uint64_t seed, int witness_like) {
    uint64_t s = seed ? seed : 0x5EEDB2000001ULL;
    for (uint64_t i = 0; i < n; ++i) {
        uint8_t rep[32];
        fr_t v;
        if (witness_like) {
            uint64_t sel = xs64star(&s) % 10;
            if (sel < 6) {
                memset(&out[i], 0, sizeof out[i]);
                continue;
            }
            if (sel < 9) {
                fr_from_u64(&out[i], xs64star(&s) & 0xffff);
                continue;
            }
        }
        for (;;) {
            uint64_t l[4] = {xs64star(&s), xs64star(&s), xs64star(&s), xs64star(&s) & 0x3fffffffffffffffULL};
            memcpy(rep, l, 32);
            if (fr_from_repr(&v, rep)) break;
        }
        out[i] = v;
    }
}

typedef struct {
    g1_affine_t *out;
    uint64_t lo, hi, seed;
} pts_job_t;

static void *pts_worker(void *arg) {
    pts_job_t *j = (pts_job_t *)arg;
    g1_affine_t gen;
    g1_generator(&gen);
    g1_t G;
    g1_from_affine(&G, &gen);
    uint64_t cnt = j->hi - j->lo;
    g1_t *tmp = (g1_t *)malloc(sizeof(g1_t) * (cnt ? cnt : 1));
    for (uint64_t i = j->lo; i < j->hi; ++i) {
        fr_t s;
        oracle_fill_fr(&s, 1, j->seed * 0x9E3779B97F4A7C15ULL + i + 1, 0);
        g1_mul(&tmp[i - j->lo], &G, &s);
    }
    g1_batch_normalize(j->out + j->lo, tmp, cnt);
    free(tmp);
    return NULL;
}

/* P_i = s_i * G with s_i drawn from the generator above (independent per index) */
void oracle_fill_points(g1_affine_t *out, uint64_t n, uint64_t seed, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    pts_job_t *jobs = (pts_job_t *)malloc(sizeof(pts_job_t) * threads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t].out = out;
        jobs[t].lo = n * t / threads;
        jobs[t].hi = n * (t + 1) / threads;
        jobs[t].seed = seed;
        pthread_create(&tids[t], NULL, pts_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
}

/* Fast bulk generator of distinct valid curve points for the large CPU-baseline runs:
 * thread chunk [lo,hi): P_lo = s_lo*G, P_{i+1} = P_i + D with one fixed D; batch-normalised. */
typedef struct {
    g1_affine_t *out;
    uint64_t lo, hi, seed;
} chain_job_t;

static void *chain_worker(void *arg) {
    chain_job_t *j = (chain_job_t *)arg;
    g1_affine_t gen, D;
    g1_generator(&gen);
    g1_t G, P, Dj;
    g1_from_affine(&G, &gen);
    fr_t s;
    oracle_fill_fr(&s, 1, j->seed ^ 0xD1FFD1FFULL, 0);
    g1_mul(&Dj, &G, &s);
    g1_to_affine(&D, &Dj);
    oracle_fill_fr(&s, 1, j->seed * 0x9E3779B97F4A7C15ULL + j->lo + 1, 0);
    g1_mul(&P, &G, &s);
    const uint64_t BATCH = 4096;
    g1_t *tmp = (g1_t *)malloc(sizeof(g1_t) * BATCH);
    for (uint64_t i = j->lo; i < j->hi; i += BATCH) {
        uint64_t cnt = (j->hi - i < BATCH) ? j->hi - i : BATCH;
        for (uint64_t t = 0; t < cnt; ++t) {
            tmp[t] = P;
            g1_add_mixed(&P, &P, &D);
        }
        g1_batch_normalize(j->out + i, tmp, cnt);
    }
    free(tmp);
    return NULL;
}

void oracle_fill_points_chain(g1_affine_t *out, uint64_t n, uint64_t seed, int threads) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    chain_job_t *jobs = (chain_job_t *)malloc(sizeof(chain_job_t) * threads);
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * threads);
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (chain_job_t){out, n * t / threads, n * (t + 1) / threads, seed};
        pthread_create(&tids[t], NULL, chain_worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(tids[t], NULL);
    free(jobs);
    free(tids);
}
