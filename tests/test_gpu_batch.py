"""Batched commitments and the grouped column pipeline (b200zk_msm_g1_batch, b200zk_run_column_jobs) against the
oracle: many columns over the same bases through ONE Pippenger pipeline must give, column by column, exactly
ParamsKZG::commit_lagrange's point (halo2_proofs/src/poly/kzg/commitment.rs @ e5ddf67; oracle/halo2_arith.c) --
the shape of the inner (zkEVM super-circuit, k = 20) proof with its several hundred columns."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu
SEED = 0x5EEDB2000777


def aff(j):
    return O.g1_to_affine(j)


@pytest.mark.parametrize("log_n,count", [(0, 3), (3, 40), (9, 37), (12, 33), (16, 19)])
def test_msm_batch_matches_oracle_column_by_column(ctx, log_n, count):
    """plain bases (< 2^16 points: per-window bucket sets) and precomputed tables (2^16: one set per column);
    counts that are not multiples of the batch size; zero, witness-like and uniform columns mixed; host and
    device-resident columns mixed in one call."""
    import torch

    n = 1 << log_n
    bases = O.fill_points_chain(n, 500 + log_n, 8)
    srs = ctx.srs_register(bases)
    cols = []
    for j in range(count):
        if j % 7 == 3:
            c = np.zeros((n, 4), np.uint64)  # an all-zero column (unused advice column)
        else:
            c = O.fill_fr(n, SEED + 31 * j + log_n, witness_like=(j % 2 == 0))
        cols.append(c)
    mixed = [torch.from_numpy(c.view(np.int64)).cuda() if j % 3 == 1 else c for j, c in enumerate(cols)]
    torch.cuda.synchronize()
    got = srs.msm_batch(mixed)
    for j, c in enumerate(cols):
        exp = O.best_multiexp(c, bases, threads=8)
        assert np.array_equal(aff(got[j]), aff(exp)), (log_n, j)
        assert np.array_equal(aff(got[j]), aff(srs.msm(c))), (log_n, j)  # and the single-column entry point
    assert srs.msm_batch([]).shape == (0, 12)
    srs.release()


def test_msm_batch_skewed_and_giant_buckets_do_not_leak_between_columns(ctx):
    """one column of equal scalars (a giant bucket, split over threads and recombined) beside ordinary columns"""
    n = 1 << 14
    bases = O.fill_points_chain(n, 901, 8)
    srs = ctx.srs_register(bases)
    giant = np.tile(O.fr_from_int(7), (n, 1))
    giant[77] = O.fr_from_int(O.R_MOD - 3)
    cols = [O.fill_fr(n, SEED + 1), giant, O.fill_fr(n, SEED + 2, witness_like=True), np.tile(O.fr_from_int(1), (n, 1))]
    got = srs.msm_batch(cols)
    for j, c in enumerate(cols):
        assert np.array_equal(aff(got[j]), aff(O.best_multiexp(c, bases, threads=8))), j
    srs.release()


def test_grouped_pipeline_many_columns_all_modes(ctx, zk):
    """run_column_jobs over more columns than one group holds: commitments (batched), coefficients, extended cosets,
    a mode-5 job (coefficients -> extended), a mode-4 job in the middle, host and device-resident inputs."""
    import torch

    k = 9
    n = 1 << k
    gl = O.fill_points_chain(n, 1201, 8)
    g = O.fill_points_chain(n, 1202, 8)
    s_gl, s_g = ctx.srs_register(gl, zk.SRS_G_LAGRANGE), ctx.srs_register(g, zk.SRS_G)
    dom, dom_o = zk.EvaluationDomain(ctx, 5, k), O.EvaluationDomain(5, k)
    ncols = 41
    cols = [O.fill_fr(n, SEED + 100 + i, witness_like=(i % 3 != 0)) for i in range(ncols)]
    coeffs = [dom_o.lagrange_to_coeff(c, threads=4) for c in cols]
    co = [torch.empty((n, 4), dtype=torch.int64, device="cuda") for _ in range(ncols)]
    eo = [torch.empty((4 * n, 4), dtype=torch.int64, device="cuda") for _ in range(ncols)]
    jobs = []
    for i, c in enumerate(cols):
        src = torch.from_numpy(c.view(np.int64)).cuda() if i % 5 == 2 else (torch.from_numpy(c.view(np.int64)).pin_memory() if i % 5 == 4 else c)
        mode = (2, 0, 1, 3, 2)[i % 5]
        jobs.append((src, s_gl if mode <= 2 else None, mode, co[i] if mode in (1, 2, 3) else None, eo[i] if mode in (2, 3) else None))
    # in the middle: the quotient's extended_to_coeff and a coefficient-form commit + a coefficients -> coset job
    ext_h = dom_o.coeff_to_extended(coeffs[1], threads=4)
    q4 = torch.empty((4 * n, 4), dtype=torch.int64, device="cuda")
    e5 = torch.empty((4 * n, 4), dtype=torch.int64, device="cuda")
    jobs.insert(20, (ext_h, None, 4, q4, None))
    jobs.insert(21, (coeffs[3], s_g, 0, None, None))
    jobs.insert(22, (coeffs[6], None, 5, None, e5))
    torch.cuda.synchronize()
    res = zk.run_column_jobs(ctx, jobs, k, omega_inv=dom.omega_inv, extended_omega=dom.extended_omega,
                             extended_omega_inv=dom.extended_omega_inv, extended_k=k + 2)
    ctx.synchronize()
    ji = 0
    for i, c in enumerate(cols):
        if ji == 20:
            ji = 23
        mode = jobs[ji][2]
        if mode <= 2:
            assert np.array_equal(aff(res[ji]), aff(O.best_multiexp(c, gl, threads=4))), i
        else:
            assert not res[ji].any()
        if mode in (1, 2, 3):
            assert np.array_equal(co[i].cpu().numpy().view(np.uint64), coeffs[i]), i
        if mode in (2, 3):
            assert np.array_equal(eo[i].cpu().numpy().view(np.uint64), dom_o.coeff_to_extended(coeffs[i], threads=4)), i
        ji += 1
    assert np.array_equal(q4.cpu().numpy().view(np.uint64), dom_o.extended_to_coeff(ext_h, threads=4))
    assert np.array_equal(aff(res[21]), aff(O.best_multiexp(coeffs[3], g, threads=4)))
    assert np.array_equal(e5.cpu().numpy().view(np.uint64), dom_o.coeff_to_extended(coeffs[6], threads=4))
    s_gl.release()
    s_g.release()
