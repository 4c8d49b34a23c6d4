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

/* ---- evaluate_h, the permutation section (plonk/evaluation.rs, `// Permutations`), for all rows of the extended domain:
 *   value = value*y + (1 - z_0[idx]) * l0[idx]
 *   value = value*y + (z_last[idx]^2 - z_last[idx]) * l_last[idx]
 *   for sets i >= 1:  value = value*y + (z_i[idx] - z_{i-1}[r_last]) * l0[idx]
 *   for every set (columns in chunks of chunk_len):
 *       left  = z_i[r_next] * prod_j (v_j[idx] + beta*sigma_j[idx] + gamma)
 *       right = z_i[idx]    * prod_j (v_j[idx] + current_delta + gamma),  current_delta *= DELTA after each column
 *       value = value*y + (left - right) * l_active_row[idx]
 * with current_delta starting at beta * ZETA * extended_omega^idx for the first column of the first set,
 * r_next = (idx + rot_scale) mod isize, r_last = (idx + last_rotation*rot_scale) mod isize. */
int halo2_permutation_h_terms(const fr_t *const *z_cosets, uint32_t n_sets, uint32_t chunk_len, const fr_t *const *value_cosets,
                              const fr_t *const *sigma_cosets, uint32_t n_cols, const fr_t *l0, const fr_t *l_last,
                              const fr_t *l_active_row, const fr_t *beta, const fr_t *gamma, const fr_t *y, const fr_t *delta,
                              const fr_t *extended_omega, int32_t last_rotation, fr_t *values, uint32_t log_size,
                              int32_t rot_scale) {
    const int64_t isize = (int64_t)1 << log_size;
    if (n_sets == 0) return 0;
    fr_t beta_term = fr_ONE, delta_start;
    fr_mul(&delta_start, beta, &fr_ZETA);
    for (int64_t idx = 0; idx < isize; ++idx) {
        int64_t r_next = (idx + rot_scale) % isize, r_last = (idx + (int64_t)last_rotation * rot_scale) % isize;
        if (r_next < 0) r_next += isize;
        if (r_last < 0) r_last += isize;
        fr_t v = values[idx], t, u;
#define FOLD(term) do { fr_mul(&v, &v, y); fr_add(&v, &v, (term)); } while (0)
        fr_sub(&t, &fr_ONE, &z_cosets[0][idx]);
        fr_mul(&t, &t, &l0[idx]);
        FOLD(&t);
        const fr_t *zl = &z_cosets[n_sets - 1][idx];
        fr_sqr(&t, zl);
        fr_sub(&t, &t, zl);
        fr_mul(&t, &t, &l_last[idx]);
        FOLD(&t);
        for (uint32_t s = 1; s < n_sets; ++s) {
            fr_sub(&t, &z_cosets[s][idx], &z_cosets[s - 1][r_last]);
            fr_mul(&t, &t, &l0[idx]);
            FOLD(&t);
        }
        fr_t current_delta;
        fr_mul(&current_delta, &delta_start, &beta_term);
        for (uint32_t s = 0; s < n_sets; ++s) {
            uint32_t c0 = s * chunk_len, c1 = c0 + chunk_len;
            if (c1 > n_cols) c1 = n_cols;
            fr_t left = z_cosets[s][r_next], right = z_cosets[s][idx];
            for (uint32_t j = c0; j < c1; ++j) {
                fr_mul(&u, beta, &sigma_cosets[j][idx]);
                fr_add(&u, &u, &value_cosets[j][idx]);
                fr_add(&u, &u, gamma);
                fr_mul(&left, &left, &u);
            }
            for (uint32_t j = c0; j < c1; ++j) {
                fr_add(&u, &value_cosets[j][idx], &current_delta);
                fr_add(&u, &u, gamma);
                fr_mul(&right, &right, &u);
                fr_mul(&current_delta, &current_delta, delta);
            }
            fr_sub(&t, &left, &right);
            fr_mul(&t, &t, &l_active_row[idx]);
            FOLD(&t);
        }
        values[idx] = v;
        fr_mul(&beta_term, &beta_term, extended_omega);
    }
    return 0;
}

/* ---- evaluate_h, one log-derivative lookup (mv_lookup): phi_i = inputs_i + beta (the compressed input expressions,
 * already evaluated on the coset), tau = table + beta:
 *   inputs_prod = prod_i phi_i;  inputs_inv_sum = sum_i 1/phi_i  (BatchInvert: zero stays zero)
 *   lhs = tau * inputs_prod * (phi[r_next] - phi[idx]);  rhs = inputs_prod * (tau * inputs_inv_sum - m[idx])
 *   value = value*y + l0*phi;  value = value*y + l_last*phi;  value = value*y + (lhs - rhs) * l_active_row */
int halo2_logup_h_terms(const fr_t *const *input_cosets, uint32_t n_inputs, const fr_t *table_coset, const fr_t *m_coset,
                        const fr_t *phi_coset, const fr_t *l0, const fr_t *l_last, const fr_t *l_active_row, const fr_t *beta,
                        const fr_t *y, fr_t *values, uint32_t log_size, int32_t rot_scale) {
    const int64_t isize = (int64_t)1 << log_size;
    fr_t *ph = (fr_t *)malloc(sizeof(fr_t) * (n_inputs ? n_inputs : 1)), *iv = (fr_t *)malloc(sizeof(fr_t) * (n_inputs ? n_inputs : 1));
    fr_t *scratch = (fr_t *)malloc(sizeof(fr_t) * (n_inputs ? n_inputs : 1));
    if (!ph || !iv || !scratch) return -1;
    for (int64_t idx = 0; idx < isize; ++idx) {
        int64_t r_next = (idx + rot_scale) % isize;
        if (r_next < 0) r_next += isize;
        fr_t prod = fr_ONE, inv_sum, tau, lhs, rhs, t, v = values[idx];
        memset(&inv_sum, 0, sizeof inv_sum);
        for (uint32_t i = 0; i < n_inputs; ++i) {
            fr_add(&ph[i], &input_cosets[i][idx], beta);
            iv[i] = ph[i];
            fr_mul(&prod, &prod, &ph[i]);
        }
        fr_batch_invert(iv, n_inputs, scratch);
        for (uint32_t i = 0; i < n_inputs; ++i) fr_add(&inv_sum, &inv_sum, &iv[i]);
        fr_add(&tau, &table_coset[idx], beta);
        fr_sub(&t, &phi_coset[r_next], &phi_coset[idx]);
        fr_mul(&lhs, &tau, &prod);
        fr_mul(&lhs, &lhs, &t);
        fr_mul(&rhs, &tau, &inv_sum);
        fr_sub(&rhs, &rhs, &m_coset[idx]);
        fr_mul(&rhs, &rhs, &prod);
        fr_mul(&t, &l0[idx], &phi_coset[idx]);
        FOLD(&t);
        fr_mul(&t, &l_last[idx], &phi_coset[idx]);
        FOLD(&t);
        fr_sub(&t, &lhs, &rhs);
        fr_mul(&t, &t, &l_active_row[idx]);
        FOLD(&t);
        values[idx] = v;
    }
#undef FOLD
    free(ph);
    free(iv);
    free(scratch);
    return 0;
}
