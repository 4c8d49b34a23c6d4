/*
 * ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or called
 * from the product path (scroll-prover_b200/). Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * 4x64-bit-limb Montgomery prime field, instantiated twice (Fr, Fq) by
 * bn254_field.c.  CPU restatement of halo2curves 0.1.0 @ scroll-tech/halo2curves
 * 112f5b9 (reference pin: /root/reference/Cargo.lock:1911-1913), files
 * src/bn256/{fr.rs,fq.rs} + src/derive/field.rs (field_arithmetic! macro):
 *   - element = [u64;4] little-endian limbs holding a*R mod p, R = 2^256
 *   - mul = schoolbook 4x4 product then montgomery_reduce; result in [0,p)
 *   - add/sub = limb add/sub then conditional -p / +p
 *   - to_repr = montgomery_reduce(a,0,0,0,0) -> 32 B LE canonical
 *   - from_repr rejects values >= p
 * The source is NOT under /root/reference (un-vendored git dependency), so this
 * follows the published algorithm; constants are pinned by
 * release-v0.13.1/chunk.protocol + evm_verifier.yul:17-18 (tests/test_oracle_fixtures.py).
 *
 * Usage: #define FP_NAME fr / FP_MOD0..3 / FP_INV then #include this file.
 */
#include <stdint.h>
#include <string.h>

#define FP_CAT_(a, b) a##_##b
#define FP_CAT(a, b) FP_CAT_(a, b)
#define FP_FN(name) FP_CAT(FP_NAME, name)
#define FP_T FP_CAT(FP_NAME, t)

typedef unsigned __int128 FP_FN(u128);

static const uint64_t FP_FN(MOD)[4] = {FP_MOD0, FP_MOD1, FP_MOD2, FP_MOD3};

/* returns 1 if a >= MOD */
static inline int FP_FN(geq_mod)(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > FP_FN(MOD)[i]) return 1;
        if (a[i] < FP_FN(MOD)[i]) return 0;
    }
    return 1;
}

static inline void FP_FN(sub_mod_raw)(uint64_t a[4]) {
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 d = (unsigned __int128)a[i] - FP_FN(MOD)[i] - (uint64_t)br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}

void FP_FN(add)(FP_T *r, const FP_T *a, const FP_T *b) {
    uint64_t t[4];
    unsigned __int128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (unsigned __int128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    /* p < 2^254 so no carry out of limb 3 for reduced inputs */
    if (FP_FN(geq_mod)(t)) FP_FN(sub_mod_raw)(t);
    memcpy(r->l, t, 32);
}

void FP_FN(sub)(FP_T *r, const FP_T *a, const FP_T *b) {
    uint64_t t[4];
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (unsigned __int128)t[i] + FP_FN(MOD)[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, 32);
}

void FP_FN(neg)(FP_T *r, const FP_T *a) {
    FP_T z;
    memset(&z, 0, sizeof z);
    FP_FN(sub)(r, &z, a);
}

void FP_FN(dbl)(FP_T *r, const FP_T *a) { FP_FN(add)(r, a, a); }

/* montgomery_reduce of an 8-limb value t (t < p*2^256): returns t * R^-1 mod p in [0,p) */
static inline void FP_FN(mont_reduce)(uint64_t r[4], uint64_t t[8]) {
    uint64_t carry2 = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t k = t[i] * (uint64_t)FP_INV;
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (unsigned __int128)k * FP_FN(MOD)[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        unsigned __int128 s = (unsigned __int128)t[i + 4] + (uint64_t)c + carry2;
        t[i + 4] = (uint64_t)s;
        carry2 = (uint64_t)(s >> 64);
    }
    uint64_t o[4] = {t[4], t[5], t[6], t[7]};
    if (carry2 || FP_FN(geq_mod)(o)) FP_FN(sub_mod_raw)(o);
    memcpy(r, o, 32);
}

void FP_FN(mul)(FP_T *r, const FP_T *a, const FP_T *b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; ++i) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (unsigned __int128)a->l[i] * b->l[j] + t[i + j];
            t[i + j] = (uint64_t)c;
            c >>= 64;
        }
        t[i + 4] = (uint64_t)c;
    }
    FP_FN(mont_reduce)(r->l, t);
}

void FP_FN(sqr)(FP_T *r, const FP_T *a) { FP_FN(mul)(r, a, a); }

/* canonical little-endian bytes (to_repr) */
void FP_FN(to_repr)(uint8_t out[32], const FP_T *a) {
    uint64_t t[8] = {a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0};
    uint64_t c[4];
    FP_FN(mont_reduce)(c, t);
    memcpy(out, c, 32); /* little-endian host */
}

/* canonical integer limbs -> Montgomery form; returns 0 if value >= p (from_repr's None) */
int FP_FN(from_repr)(FP_T *r, const uint8_t in[32]) {
    FP_T t;
    memcpy(t.l, in, 32);
    if (FP_FN(geq_mod)(t.l)) return 0;
    FP_FN(mul)(r, &t, &FP_FN(R2));
    return 1;
}

int FP_FN(is_zero)(const FP_T *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
int FP_FN(eq)(const FP_T *a, const FP_T *b) { return memcmp(a->l, b->l, 32) == 0; }

/* pow by a 4-limb little-endian exponent (ff::Field::pow_vartime, square-and-multiply MSB first) */
void FP_FN(pow)(FP_T *r, const FP_T *a, const uint64_t e[4]) {
    FP_T acc = FP_FN(ONE);
    for (int i = 3; i >= 0; --i)
        for (int b = 63; b >= 0; --b) {
            FP_FN(sqr)(&acc, &acc);
            if ((e[i] >> b) & 1) FP_FN(mul)(&acc, &acc, a);
        }
    *r = acc;
}

/* invert = a^(p-2) (halo2curves uses an addition chain; same value). 0 -> 0 with return 0. */
int FP_FN(inv)(FP_T *r, const FP_T *a) {
    if (FP_FN(is_zero)(a)) {
        memset(r, 0, sizeof *r);
        return 0;
    }
    uint64_t e[4] = {FP_FN(MOD)[0] - 2, FP_FN(MOD)[1], FP_FN(MOD)[2], FP_FN(MOD)[3]};
    FP_FN(pow)(r, a, e);
    return 1;
}

/* ff::BatchInvert semantics: zeros are skipped and left as zero */
void FP_FN(batch_invert)(FP_T *v, uint64_t n, FP_T *scratch) {
    FP_T acc = FP_FN(ONE);
    for (uint64_t i = 0; i < n; ++i) {
        scratch[i] = acc;
        if (!FP_FN(is_zero)(&v[i])) FP_FN(mul)(&acc, &acc, &v[i]);
    }
    FP_FN(inv)(&acc, &acc);
    for (uint64_t i = n; i-- > 0;) {
        if (FP_FN(is_zero)(&v[i])) continue;
        FP_T t;
        FP_FN(mul)(&t, &scratch[i], &acc);
        FP_FN(mul)(&acc, &acc, &v[i]);
        v[i] = t;
    }
}

void FP_FN(from_u64)(FP_T *r, uint64_t v) {
    FP_T t = {{v, 0, 0, 0}};
    FP_FN(mul)(r, &t, &FP_FN(R2));
}

#undef FP_NAME
#undef FP_MOD0
#undef FP_MOD1
#undef FP_MOD2
#undef FP_MOD3
#undef FP_INV
