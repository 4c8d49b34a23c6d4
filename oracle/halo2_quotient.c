/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of the prover work between the transforms of plonk::create_proof:
 *   - plonk::evaluation::GraphEvaluator::evaluate / Calculation::evaluate     (evaluation.rs)
 *   - permutation::Argument::commit, one iteration of its column-chunk loop    (permutation/prover.rs)
 *   - the phi(X) running sum of the log-derivative lookup                       (mv_lookup/prover.rs)
 * of halo2_proofs 1.1.0 @ scroll-tech/halo2 e5ddf67 (pin /root/reference/Cargo.lock:1886-1888; un-vendored, restated
 * from the published algorithm -- SURVEY.md Appendix A conventions; entered from
 * /root/reference/integration/src/prove.rs:37-39).
 *
 * PARITY STATUS: "parity unpinned" -- the reference ships no input/output pairs for these steps.  The restatement
 * is cross-checked against an independent big-integer model (tests/test_quotient_oracle.py) and against algebraic
 * identities (z telescopes to 1 for a valid permutation, phi telescopes to 0 for a valid lookup).
 */
#include <stdlib.h>
#include <string.h>

#include "bn254_oracle.h"

/* GraphEvaluator::evaluate for every row idx < 2^log_size:
 *   rotations[r] -> row (idx + rotations[r] * rot_scale) mod isize          (get_rotation_idx)
 *   intermediates[i] = calculations[i].evaluate(...)                         (one per calculation, in order)
 *   values[idx] = intermediates[last]  (0 when there are no calculations), PreviousValue = the old values[idx]
 * ExtendedX (not an upstream ValueSource; the test programs use it for the permutation identity) is
 * g_coset * extended_omega^idx with g_coset = ZETA, i.e. the idx-th point of the extended coset. */
static const fr_t *gq_get(const halo2_value_source_t *s, const fr_t *constants, const fr_t *intermediates,
                          const fr_t *const *fixed, const fr_t *const *advice, const fr_t *const *instance,
                          const fr_t *challenges, const fr_t *beta, const fr_t *gamma, const fr_t *theta, const fr_t *y,
                          const fr_t *previous, const fr_t *ext_x, const uint64_t *rot_rows) {
    switch (s->kind) {
        case HALO2_SRC_CONSTANT: return &constants[s->index];
        case HALO2_SRC_INTERMEDIATE: return &intermediates[s->index];
        case HALO2_SRC_FIXED: return &fixed[s->index][rot_rows[s->rotation]];
        case HALO2_SRC_ADVICE: return &advice[s->index][rot_rows[s->rotation]];
        case HALO2_SRC_INSTANCE: return &instance[s->index][rot_rows[s->rotation]];
        case HALO2_SRC_CHALLENGE: return &challenges[s->index];
        case HALO2_SRC_BETA: return beta;
        case HALO2_SRC_GAMMA: return gamma;
        case HALO2_SRC_THETA: return theta;
        case HALO2_SRC_Y: return y;
        case HALO2_SRC_PREVIOUS_VALUE: return previous;
        default: return ext_x;
    }
}

int halo2_graph_evaluate(const halo2_calculation_t *calcs, uint32_t n_calcs, const halo2_value_source_t *parts,
                         const fr_t *constants, const int32_t *rotations, uint32_t n_rotations, const fr_t *const *fixed,
                         const fr_t *const *advice, const fr_t *const *instance, const fr_t *challenges, const fr_t *beta,
                         const fr_t *gamma, const fr_t *theta, const fr_t *y, const fr_t *extended_omega, fr_t *values,
                         uint32_t log_size, int32_t rot_scale) {
    const int64_t isize = (int64_t)1 << log_size;
    fr_t *inter = (fr_t *)malloc(sizeof(fr_t) * (n_calcs ? n_calcs : 1));
    uint64_t *rot_rows = (uint64_t *)malloc(sizeof(uint64_t) * (n_rotations ? n_rotations : 1));
    if (!inter || !rot_rows) return -1;
    fr_t x = fr_ZETA; /* ZETA * extended_omega^idx, stepped */
#define GET(src) gq_get(src, constants, inter, fixed, advice, instance, challenges, beta, gamma, theta, y, &prev, &x, rot_rows)
    for (int64_t idx = 0; idx < isize; ++idx) {
        for (uint32_t r = 0; r < n_rotations; ++r) {
            int64_t v = (idx + (int64_t)rotations[r] * rot_scale) % isize; /* rem_euclid */
            if (v < 0) v += isize;
            rot_rows[r] = (uint64_t)v;
        }
        fr_t prev = values[idx];
        for (uint32_t i = 0; i < n_calcs; ++i) {
            const halo2_calculation_t *c = &calcs[i];
            fr_t r;
            switch (c->op) {
                case HALO2_CALC_ADD: fr_add(&r, GET(&c->a), GET(&c->b)); break;
                case HALO2_CALC_SUB: fr_sub(&r, GET(&c->a), GET(&c->b)); break;
                case HALO2_CALC_MUL: fr_mul(&r, GET(&c->a), GET(&c->b)); break;
                case HALO2_CALC_SQUARE: fr_sqr(&r, GET(&c->a)); break;
                case HALO2_CALC_DOUBLE: fr_dbl(&r, GET(&c->a)); break;
                case HALO2_CALC_NEGATE: fr_neg(&r, GET(&c->a)); break;
                case HALO2_CALC_HORNER: {
                    const fr_t *factor = GET(&c->b);
                    r = *GET(&c->a);
                    for (uint32_t j = 0; j < c->parts_len; ++j) {
                        fr_mul(&r, &r, factor);
                        fr_add(&r, &r, GET(&parts[c->parts_offset + j]));
                    }
                    break;
                }
                default: r = *GET(&c->a); break; /* Store */
            }
            inter[i] = r;
        }
        if (n_calcs) values[idx] = inter[n_calcs - 1];
        else memset(&values[idx], 0, sizeof(fr_t));
        if (extended_omega) fr_mul(&x, &x, extended_omega);
    }
#undef GET
    free(inter);
    free(rot_rows);
    return 0;
}

/* out[0] = init; out[i] = out[i-1] (*|+) in[i-1] */
void halo2_prefix_scan(int op, const fr_t *in, uint64_t n, const fr_t *init, fr_t *out) {
    fr_t acc = *init;
    for (uint64_t i = 0; i < n; ++i) {
        fr_t v = in[i];
        out[i] = acc;
        if (op == 0) fr_mul(&acc, &acc, &v);
        else fr_add(&acc, &acc, &v);
    }
}

/* permutation::Argument::commit, the body of `for (columns, permutations) in chunks`:
 * modified_values = 1; *= beta*sigma + gamma + value per column; batch_invert; then per column
 * *= deltaomega*beta + gamma + value with deltaomega = delta^(column position) * omega^row (stepped by omega along the
 * rows and by DELTA from one column to the next); z[0] = last_z; z[row] = z[row-1] * modified_values[row-1]. */
int halo2_permutation_product(const fr_t *const *values, const fr_t *const *sigma, uint32_t n_cols, const fr_t *beta,
                              const fr_t *gamma, const fr_t *delta_omega_start, const fr_t *delta, const fr_t *omega,
                              uint32_t k, const fr_t *z_init, fr_t *z_out) {
    const uint64_t n = (uint64_t)1 << k;
    fr_t *mv = (fr_t *)malloc(sizeof(fr_t) * n), *scratch = (fr_t *)malloc(sizeof(fr_t) * n);
    if (!mv || !scratch) return -1;
    for (uint64_t i = 0; i < n; ++i) mv[i] = fr_ONE;
    for (uint32_t j = 0; j < n_cols; ++j)
        for (uint64_t i = 0; i < n; ++i) {
            fr_t t;
            fr_mul(&t, beta, &sigma[j][i]);
            fr_add(&t, &t, gamma);
            fr_add(&t, &t, &values[j][i]);
            fr_mul(&mv[i], &mv[i], &t);
        }
    fr_batch_invert(mv, n, scratch);
    fr_t deltaomega0 = *delta_omega_start;
    for (uint32_t j = 0; j < n_cols; ++j) {
        fr_t deltaomega = deltaomega0;
        for (uint64_t i = 0; i < n; ++i) {
            fr_t t;
            fr_mul(&t, &deltaomega, beta);
            fr_add(&t, &t, gamma);
            fr_add(&t, &t, &values[j][i]);
            fr_mul(&mv[i], &mv[i], &t);
            fr_mul(&deltaomega, &deltaomega, omega);
        }
        fr_mul(&deltaomega0, &deltaomega0, delta);
    }
    halo2_prefix_scan(0, mv, n, z_init, z_out);
    free(mv);
    free(scratch);
    return 0;
}

/* log-derivative lookup: inputs_log_derivatives[i] = sum_j 1/(f_j[i] + beta); table_log_derivatives[i] =
 * m[i]/(t[i] + beta) (both through BatchInvert: a zero denominator stays zero); phi[0] = phi_init;
 * phi[i] = phi[i-1] + inputs_log_derivatives[i-1] - table_log_derivatives[i-1]. */
int halo2_logup_running_sum(const fr_t *const *inputs, uint32_t n_inputs, const fr_t *table, const fr_t *m,
                            const fr_t *beta, uint32_t k, const fr_t *phi_init, fr_t *phi_out) {
    const uint64_t n = (uint64_t)1 << k;
    fr_t *den = (fr_t *)malloc(sizeof(fr_t) * n), *scratch = (fr_t *)malloc(sizeof(fr_t) * n);
    fr_t *d = (fr_t *)calloc(n, sizeof(fr_t));
    if (!den || !scratch || !d) return -1;
    for (uint32_t j = 0; j < n_inputs; ++j) {
        for (uint64_t i = 0; i < n; ++i) fr_add(&den[i], &inputs[j][i], beta);
        fr_batch_invert(den, n, scratch);
        for (uint64_t i = 0; i < n; ++i) fr_add(&d[i], &d[i], &den[i]);
    }
    for (uint64_t i = 0; i < n; ++i) fr_add(&den[i], &table[i], beta);
    fr_batch_invert(den, n, scratch);
    for (uint64_t i = 0; i < n; ++i) {
        fr_t t;
        fr_mul(&t, &den[i], &m[i]);
        fr_sub(&d[i], &d[i], &t);
    }
    halo2_prefix_scan(1, d, n, phi_init, phi_out);
    free(den);
    free(scratch);
    free(d);
    return 0;
}
