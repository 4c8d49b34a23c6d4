/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 * BN254 G1: y^2 = x^3 + 3 over Fq, generator (1,2), cofactor 1.
 * Restates halo2curves 0.1.0 src/bn256/curve.rs + src/derive/curve.rs
 * (new_curve_impl!: Jacobian add / mixed add / double with a = 0, to_affine,
 * batch_normalize, compressed encoding) — pin /root/reference/Cargo.lock:1911-1913;
 * type used by the reference at integration/src/prove.rs:1.
 * Group-element results are representation-independent, so only the affine
 * (normalised) coordinates are ever compared.
 */
#include <stdlib.h>
#include <string.h>

#include "bn254_oracle.h"

static const fq_t FQ_B = {{0x7a17caa950ad28d7ULL, 0x1f6ac17ae15521b9ULL, 0x334bea4e696bd284ULL, 0x2a1f6744ce179d8eULL}}; /* 3 */

void g1_identity(g1_t *r) {
    memset(r, 0, sizeof *r);
    r->y = fq_ONE; /* halo2curves identity = (0, 1, 0) */
}
int g1_is_identity(const g1_t *p) { return fq_is_zero(&p->z); }
int g1_affine_is_identity(const g1_affine_t *p) { return fq_is_zero(&p->x) && fq_is_zero(&p->y); }

void g1_from_affine(g1_t *r, const g1_affine_t *p) {
    if (g1_affine_is_identity(p)) {
        g1_identity(r);
        return;
    }
    r->x = p->x;
    r->y = p->y;
    r->z = fq_ONE;
}

int g1_affine_is_on_curve(const g1_affine_t *p) {
    if (g1_affine_is_identity(p)) return 1;
    fq_t y2, x3;
    fq_sqr(&y2, &p->y);
    fq_sqr(&x3, &p->x);
    fq_mul(&x3, &x3, &p->x);
    fq_add(&x3, &x3, &FQ_B);
    return fq_eq(&y2, &x3);
}

void g1_affine_neg(g1_affine_t *r, const g1_affine_t *p) {
    r->x = p->x;
    fq_neg(&r->y, &p->y);
}

void g1_neg(g1_t *r, const g1_t *p) {
    r->x = p->x;
    fq_neg(&r->y, &p->y);
    r->z = p->z;
}

/* dbl-2009-l (a = 0) */
void g1_double(g1_t *r, const g1_t *p) {
    if (g1_is_identity(p)) {
        g1_identity(r);
        return;
    }
    fq_t a, b, c, d, e, f, t, x3, y3, z3;
    fq_sqr(&a, &p->x);
    fq_sqr(&b, &p->y);
    fq_sqr(&c, &b);
    fq_add(&d, &p->x, &b);
    fq_sqr(&d, &d);
    fq_sub(&d, &d, &a);
    fq_sub(&d, &d, &c);
    fq_dbl(&d, &d);
    fq_dbl(&e, &a);
    fq_add(&e, &e, &a);
    fq_sqr(&f, &e);
    fq_mul(&z3, &p->z, &p->y);
    fq_dbl(&z3, &z3);
    fq_dbl(&t, &d);
    fq_sub(&x3, &f, &t);
    fq_dbl(&c, &c);
    fq_dbl(&c, &c);
    fq_dbl(&c, &c);
    fq_sub(&t, &d, &x3);
    fq_mul(&y3, &e, &t);
    fq_sub(&y3, &y3, &c);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}

/* add-2007-bl with identity / doubling / inverse handling */
void g1_add(g1_t *r, const g1_t *p, const g1_t *q) {
    if (g1_is_identity(p)) {
        *r = *q;
        return;
    }
    if (g1_is_identity(q)) {
        *r = *p;
        return;
    }
    fq_t z1z1, z2z2, u1, u2, s1, s2, t;
    fq_sqr(&z1z1, &p->z);
    fq_sqr(&z2z2, &q->z);
    fq_mul(&u1, &p->x, &z2z2);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s1, &p->y, &z2z2);
    fq_mul(&s1, &s1, &q->z);
    fq_mul(&s2, &q->y, &z1z1);
    fq_mul(&s2, &s2, &p->z);
    if (fq_eq(&u1, &u2)) {
        if (fq_eq(&s1, &s2))
            g1_double(r, p);
        else
            g1_identity(r);
        return;
    }
    fq_t h, i, j, rr, v, x3, y3, z3;
    fq_sub(&h, &u2, &u1);
    fq_dbl(&i, &h);
    fq_sqr(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&rr, &s2, &s1);
    fq_dbl(&rr, &rr);
    fq_mul(&v, &u1, &i);
    fq_sqr(&x3, &rr);
    fq_sub(&x3, &x3, &j);
    fq_sub(&x3, &x3, &v);
    fq_sub(&x3, &x3, &v);
    fq_mul(&s1, &s1, &j);
    fq_dbl(&s1, &s1);
    fq_sub(&t, &v, &x3);
    fq_mul(&y3, &rr, &t);
    fq_sub(&y3, &y3, &s1);
    fq_add(&z3, &p->z, &q->z);
    fq_sqr(&z3, &z3);
    fq_sub(&z3, &z3, &z1z1);
    fq_sub(&z3, &z3, &z2z2);
    fq_mul(&z3, &z3, &h);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}

/* madd-2007-bl */
void g1_add_mixed(g1_t *r, const g1_t *p, const g1_affine_t *q) {
    if (g1_affine_is_identity(q)) {
        *r = *p;
        return;
    }
    if (g1_is_identity(p)) {
        g1_from_affine(r, q);
        return;
    }
    fq_t z1z1, u2, s2;
    fq_sqr(&z1z1, &p->z);
    fq_mul(&u2, &q->x, &z1z1);
    fq_mul(&s2, &q->y, &z1z1);
    fq_mul(&s2, &s2, &p->z);
    if (fq_eq(&p->x, &u2)) {
        if (fq_eq(&p->y, &s2))
            g1_double(r, p);
        else
            g1_identity(r);
        return;
    }
    fq_t h, hh, i, j, rr, v, x3, y3, z3, t;
    fq_sub(&h, &u2, &p->x);
    fq_sqr(&hh, &h);
    fq_dbl(&i, &hh);
    fq_dbl(&i, &i);
    fq_mul(&j, &h, &i);
    fq_sub(&rr, &s2, &p->y);
    fq_dbl(&rr, &rr);
    fq_mul(&v, &p->x, &i);
    fq_sqr(&x3, &rr);
    fq_sub(&x3, &x3, &j);
    fq_sub(&x3, &x3, &v);
    fq_sub(&x3, &x3, &v);
    fq_mul(&j, &p->y, &j);
    fq_dbl(&j, &j);
    fq_sub(&t, &v, &x3);
    fq_mul(&y3, &rr, &t);
    fq_sub(&y3, &y3, &j);
    fq_add(&z3, &p->z, &h);
    fq_sqr(&z3, &z3);
    fq_sub(&z3, &z3, &z1z1);
    fq_sub(&z3, &z3, &hh);
    r->x = x3;
    r->y = y3;
    r->z = z3;
}

void g1_to_affine(g1_affine_t *r, const g1_t *p) {
    if (g1_is_identity(p)) {
        memset(r, 0, sizeof *r);
        return;
    }
    fq_t zi, zi2, zi3;
    fq_inv(&zi, &p->z);
    fq_sqr(&zi2, &zi);
    fq_mul(&zi3, &zi2, &zi);
    fq_mul(&r->x, &p->x, &zi2);
    fq_mul(&r->y, &p->y, &zi3);
}

/* Curve::batch_normalize: one inversion for all z != 0 */
void g1_batch_normalize(g1_affine_t *out, const g1_t *in, uint64_t n) {
    fq_t *z = (fq_t *)malloc(sizeof(fq_t) * (n ? n : 1));
    fq_t *scratch = (fq_t *)malloc(sizeof(fq_t) * (n ? n : 1));
    for (uint64_t i = 0; i < n; ++i) z[i] = in[i].z;
    fq_batch_invert(z, n, scratch);
    for (uint64_t i = 0; i < n; ++i) {
        if (g1_is_identity(&in[i])) {
            memset(&out[i], 0, sizeof out[i]);
            continue;
        }
        fq_t zi2, zi3;
        fq_sqr(&zi2, &z[i]);
        fq_mul(&zi3, &zi2, &z[i]);
        fq_mul(&out[i].x, &in[i].x, &zi2);
        fq_mul(&out[i].y, &in[i].y, &zi3);
    }
    free(z);
    free(scratch);
}

int g1_eq(const g1_t *a, const g1_t *b) {
    int ia = g1_is_identity(a), ib = g1_is_identity(b);
    if (ia || ib) return ia && ib;
    fq_t z1z1, z2z2, l, r;
    fq_sqr(&z1z1, &a->z);
    fq_sqr(&z2z2, &b->z);
    fq_mul(&l, &a->x, &z2z2);
    fq_mul(&r, &b->x, &z1z1);
    if (!fq_eq(&l, &r)) return 0;
    fq_mul(&l, &a->y, &z2z2);
    fq_mul(&l, &l, &b->z);
    fq_mul(&r, &b->y, &z1z1);
    fq_mul(&r, &r, &a->z);
    return fq_eq(&l, &r);
}

/* scalar mul: plain MSB-first double-and-add over the canonical bits of s */
void g1_mul(g1_t *r, const g1_t *p, const fr_t *s) {
    uint8_t rep[32];
    fr_to_repr(rep, s);
    g1_t acc;
    g1_identity(&acc);
    for (int i = 255; i >= 0; --i) {
        g1_double(&acc, &acc);
        if ((rep[i >> 3] >> (i & 7)) & 1) g1_add(&acc, &acc, p);
    }
    *r = acc;
}

void g1_generator(g1_affine_t *r) {
    fq_from_u64(&r->x, 1);
    fq_from_u64(&r->y, 2);
}

void g1_affine_to_compressed(uint8_t out[32], const g1_affine_t *p) {
    if (g1_affine_is_identity(p)) {
        memset(out, 0, 32);
        out[31] |= 0x80;
        return;
    }
    uint8_t yb[32];
    fq_to_repr(out, &p->x);
    fq_to_repr(yb, &p->y);
    out[31] |= (uint8_t)((yb[0] & 1) << 6);
}

/* sqrt in Fq: q = 3 mod 4 => y = a^((q+1)/4) */
static int fq_sqrt(fq_t *r, const fq_t *a) {
    static const uint64_t e[4] = {0x4f082305b61f3f52ULL, 0x65e05aa45a1c72a3ULL, 0x6e14116da0605617ULL, 0x0c19139cb84c680aULL};
    fq_t y, y2;
    fq_pow(&y, a, e);
    fq_sqr(&y2, &y);
    if (!fq_eq(&y2, a)) return 0;
    *r = y;
    return 1;
}

int g1_affine_from_compressed(g1_affine_t *r, const uint8_t in[32]) {
    uint8_t tmp[32];
    memcpy(tmp, in, 32);
    int inf = (tmp[31] >> 7) & 1, sign = (tmp[31] >> 6) & 1;
    tmp[31] &= 0x3f;
    if (inf) {
        for (int i = 0; i < 32; ++i)
            if (tmp[i]) return 0;
        if (sign) return 0;
        memset(r, 0, sizeof *r);
        return 1;
    }
    fq_t x, y, rhs;
    if (!fq_from_repr(&x, tmp)) return 0;
    fq_sqr(&rhs, &x);
    fq_mul(&rhs, &rhs, &x);
    fq_add(&rhs, &rhs, &FQ_B);
    if (!fq_sqrt(&y, &rhs)) return 0;
    uint8_t yb[32];
    fq_to_repr(yb, &y);
    if ((yb[0] & 1) != sign) fq_neg(&y, &y);
    r->x = x;
    r->y = y;
    return 1;
}
