// Batched-affine bucket accumulation (EXPERIMENTAL; see msm_affine.cuh for the algorithm and its status).
//
// Host orchestration + thin kernel wrappers.  Replaces steps 4-5 of msm.cu's pipeline (msm_accumulate + msm_combine_*) when
// the context knob `msm_affine` is set (B200ZK_MSM_AFFINE=1); the sorted entries / bucket offsets it consumes and the XYZZ
// bucket array it fills are the ones of msm_run, so everything before and after is shared.  Not enabled by default: the
// per-thread bodies are validated on the CPU (tests/test_msm_affine_host.py), the kernels have not been timed on a B200 yet.
#include "common.cuh"
#include "msm_affine_kernels.cuh"

namespace b200zk {

static int32_t ba_invert_totals(b200zk_ctx* ctx, Fq* data, uint64_t n, Fq* scratch, const uint32_t* active) {
    if (n <= BAI_LEAF) {
        bai_leaf<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(data, n, active);
        B2_LAUNCH_CHECK(ctx);
        return B200ZK_OK;
    }
    uint64_t T = (n + BAI_SLICE - 1) / BAI_SLICE;
    Fq *prefix = scratch, *totals = scratch + n;
    uint32_t blocks = (uint32_t)((T + 255) / 256);
    bai_up<<<blocks, 256, 0, ctx->stream>>>(data, prefix, totals, n, (uint32_t)T, active);
    B2_LAUNCH_CHECK(ctx);
    B2_TRY(ba_invert_totals(ctx, totals, T, totals + T, active));
    bai_down<<<blocks, 256, 0, ctx->stream>>>(data, prefix, totals, n, (uint32_t)T, active);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}
static size_t ba_invert_scratch_elems(uint64_t n) {
    size_t e = 0;
    for (uint64_t m = n; m > BAI_LEAF;) {
        uint64_t T = (m + BAI_SLICE - 1) / BAI_SLICE;
        e += m + T;
        m = T;
    }
    return e + 1;
}

// Sums every bucket of the sorted entry list into `buckets` (XYZZ, identity for empty buckets).
//   bases / entries / offsets: as msm_run hands them to msm_accumulate (offsets has NB + 1 entries, offsets[NB] = M)
//   max_entries: host-side upper bound of M (n * W)
// Scratch (ctx->msm_affine_work): two point arrays of max_entries/2 + NB, the prefix array, the
// thread totals with their inversion scratch, two offset arrays and the scan tiles.
int32_t msm_affine_accumulate(b200zk_ctx* ctx, const Affine* bases, const uint32_t* entries, const uint32_t* offsets, uint64_t NB,
                              uint64_t max_entries, XYZZ* buckets) {
    const uint32_t BA_L = ctx->msm_affine_l ? ctx->msm_affine_l : BA_L_DEFAULT;
    const uint64_t cap1 = max_entries / 2 + NB + 1;  // outputs of the first level: sum_b ceil(m_b / 2) <= M/2 + NB
    const uint64_t tcap = (cap1 + BA_L - 1) / BA_L;
    const uint32_t ntiles = (uint32_t)((NB + BA_SCAN_TILE - 1) / BA_SCAN_TILE);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
    size_t o_p0 = carve(sizeof(Affine) * cap1), o_p1 = carve(sizeof(Affine) * cap1), o_prefix = carve(sizeof(Fq) * cap1);
    size_t o_tot = carve(sizeof(Fq) * (tcap + 1)), o_inv = carve(sizeof(Fq) * ba_invert_scratch_elems(tcap));
    size_t o_off0 = carve(4 * (NB + 1)), o_off1 = carve(4 * (NB + 1)), o_tiles = carve(4 * ((size_t)ntiles + 1)), o_flag = carve(256);
    B2_TRY(scratch_reserve(ctx, ctx->msm_affine_work, off));
    char* base = (char*)ctx->msm_affine_work.p;
    Affine* pts[2] = {(Affine*)(base + o_p0), (Affine*)(base + o_p1)};
    uint32_t* offs[2] = {(uint32_t*)(base + o_off0), (uint32_t*)(base + o_off1)};
    Fq *prefix = (Fq*)(base + o_prefix), *totals = (Fq*)(base + o_tot), *inv_scratch = (Fq*)(base + o_inv);
    uint32_t* tiles = (uint32_t*)(base + o_tiles);
    uint32_t* active = (uint32_t*)(base + o_flag);
    cudaStream_t st = ctx->stream;

    // a bucket of m entries needs ceil(log2 m) levels; m <= max_entries.  Levels past a bucket's last addition only copy its
    // single point, so running the worst-case count is correct (and cheap: the arrays shrink geometrically to <= NB points).
    uint32_t levels = 0;
    while ((1ull << levels) < max_entries) ++levels;
    BaLevel lv;
    lv.bases = bases;
    lv.entries = entries;
    lv.points = nullptr;
    lv.off_in = offsets;
    lv.NB = NB;
    uint64_t bound = max_entries;  // upper bound of the current level's input count
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t* off_out = offs[l & 1];
        Affine* out = pts[l & 1];
        ba_next_counts_tiles<<<ntiles, BA_SCAN_TPB, 0, st>>>(lv.off_in, NB, tiles);
        B2_LAUNCH_CHECK(ctx);
        ba_scan_tiles<<<1, 32, 0, st>>>(tiles, ntiles, off_out + NB, lv.off_in + NB, active);
        B2_LAUNCH_CHECK(ctx);
        ba_next_offsets<<<ntiles, BA_SCAN_TPB, 0, st>>>(lv.off_in, NB, tiles, off_out);
        B2_LAUNCH_CHECK(ctx);
        lv.off_out = off_out;
        uint64_t out_bound = bound / 2 + NB + 1;
        if (out_bound > cap1) out_bound = cap1;
        uint64_t nthreads = (out_bound + BA_L - 1) / BA_L;
        uint32_t blocks = (uint32_t)((nthreads + 127) / 128);
        ba_pass_a<<<blocks, 128, 0, st>>>(lv, prefix, totals, nthreads, BA_L, active);
        B2_LAUNCH_CHECK(ctx);
        // (threads past the real output count write a total of one, so the shared inversion stays well defined)
        B2_TRY(ba_invert_totals(ctx, totals, nthreads, inv_scratch, active));
        ba_pass_b<<<blocks, 128, 0, st>>>(lv, prefix, totals, out, nthreads, BA_L);
        B2_LAUNCH_CHECK(ctx);
        lv.entries = nullptr;
        lv.points = out;
        lv.off_in = off_out;
        bound = out_bound;
    }
    ba_finalize<<<(uint32_t)((NB + 255) / 256), 256, 0, st>>>(lv, buckets);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

}  // namespace b200zk
