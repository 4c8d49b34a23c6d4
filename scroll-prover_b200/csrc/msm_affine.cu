// Batched-affine bucket accumulation (EXPERIMENTAL; see msm_affine.cuh for the algorithm and its status).
//
// Host orchestration + thin kernel wrappers.  Replaces steps 4-5 of msm.cu's pipeline (msm_accumulate + msm_combine_*) when
// the context knob `msm_affine` is set (B200ZK_MSM_AFFINE=1); the sorted entries / bucket offsets it consumes and the XYZZ
// bucket array it fills are the ones of msm_run, so everything before and after is shared.  Not enabled by default: the
// per-thread bodies are validated on the CPU (tests/test_msm_affine_host.py), the kernels have not been timed on a B200 yet.
#include "common.cuh"
#include "msm_affine.cuh"

namespace b200zk {

constexpr uint32_t BA_L_DEFAULT = 16;  // output slots per thread: the inversion is shared by L additions per thread, then by the
                                       // hierarchical inversion of the thread totals (B200ZK_AFFINE_L overrides, for tuning)

// `active` (written by ba_scan_tiles): 0 when the level has no pair left to add -- it only copies single points, so the
// prefix products and the inversion are skipped (pass B never reads them for copies) and the level costs a few empty launches
__global__ void __launch_bounds__(128) ba_pass_a(BaLevel lv, Fq* prefix, Fq* totals, uint64_t nthreads, uint32_t L, const uint32_t* active) {
    if (!*active) return;
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) ba_thread_a(t, L, lv, prefix, totals);
}
__global__ void __launch_bounds__(128) ba_pass_b(BaLevel lv, const Fq* prefix, const Fq* inv_totals, Affine* out, uint64_t nthreads,
                                                 uint32_t L) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nthreads) ba_thread_b(t, L, lv, prefix, inv_totals, out);
}
__global__ void __launch_bounds__(256) ba_finalize(BaLevel lv, XYZZ* buckets) {
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < lv.NB) buckets[b] = ba_final_bucket(lv, b);
}

// counts of the next level, then their exclusive scan (three small kernels, NB + 1 entries)
constexpr uint32_t BA_SCAN_TPB = 256, BA_SCAN_ITEMS = 8, BA_SCAN_TILE = BA_SCAN_TPB * BA_SCAN_ITEMS;
__global__ void __launch_bounds__(BA_SCAN_TPB) ba_next_counts_tiles(const uint32_t* off_in, uint64_t NB, uint32_t* tile_sums) {
    __shared__ uint32_t sh[BA_SCAN_TPB];
    uint64_t base = (uint64_t)blockIdx.x * BA_SCAN_TILE + (uint64_t)threadIdx.x * BA_SCAN_ITEMS;
    uint32_t s = 0;
    for (uint32_t k = 0; k < BA_SCAN_ITEMS; ++k)
        if (base + k < NB) s += (off_in[base + k + 1] - off_in[base + k] + 1) >> 1;
    sh[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t d = BA_SCAN_TPB >> 1; d > 0; d >>= 1) {
        if (threadIdx.x < d) sh[threadIdx.x] += sh[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = sh[0];
}
__global__ void ba_scan_tiles(uint32_t* tile_sums, uint32_t ntiles, uint32_t* grand_total, const uint32_t* in_total, uint32_t* active) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < ntiles; ++i) {
        uint32_t v = tile_sums[i];
        tile_sums[i] = acc;
        acc += v;
    }
    *grand_total = acc;  // off_out[NB]
    *active = acc < *in_total;  // some bucket still holds two or more points
}
__global__ void __launch_bounds__(BA_SCAN_TPB) ba_next_offsets(const uint32_t* off_in, uint64_t NB, const uint32_t* tile_offs, uint32_t* off_out) {
    __shared__ uint32_t sh[BA_SCAN_TPB];
    uint64_t base = (uint64_t)blockIdx.x * BA_SCAN_TILE + (uint64_t)threadIdx.x * BA_SCAN_ITEMS;
    uint32_t v[BA_SCAN_ITEMS], s = 0;
    for (uint32_t k = 0; k < BA_SCAN_ITEMS; ++k) {
        v[k] = (base + k < NB) ? ((off_in[base + k + 1] - off_in[base + k] + 1) >> 1) : 0;
        s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (uint32_t d = 1; d < BA_SCAN_TPB; d <<= 1) {  // Hillis-Steele inclusive scan of the thread sums
        uint32_t x = sh[threadIdx.x];
        if (threadIdx.x >= d) x += sh[threadIdx.x - d];
        __syncthreads();
        sh[threadIdx.x] = x;
        __syncthreads();
    }
    uint32_t off = tile_offs[blockIdx.x] + (threadIdx.x ? sh[threadIdx.x - 1] : 0);
    for (uint32_t k = 0; k < BA_SCAN_ITEMS; ++k) {
        if (base + k < NB) off_out[base + k] = off;
        off += v[k];
    }
}

// ---- inversion of the thread totals (Fq): the same hierarchical Montgomery trick as poly.cu's batch_invert, on Fq
constexpr uint64_t BAI_SLICE = 64, BAI_LEAF = 2048;
__global__ void __launch_bounds__(256) bai_up(const Fq* data, Fq* prefix, Fq* totals, uint64_t n, uint32_t T, const uint32_t* active) {
    if (!*active) return;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    Fq acc = Fq::one();
    for (uint64_t i = t; i < n; i += T) {
        prefix[i] = acc;
        acc = acc * data[i];  // thread totals are products of non-zero denominators: never zero
    }
    totals[t] = acc;
}
__global__ void __launch_bounds__(256) bai_down(Fq* data, const Fq* prefix, const Fq* inv_totals, uint64_t n, uint32_t T, const uint32_t* active) {
    if (!*active) return;
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || t >= n) return;
    uint64_t cnt = (n - t + T - 1) / T;
    Fq acc = inv_totals[t];
    for (uint64_t j = cnt; j-- > 0;) {
        uint64_t i = t + j * T;
        Fq v = data[i];
        data[i] = prefix[i] * acc;
        acc = acc * v;
    }
}
__global__ void __launch_bounds__(128) bai_leaf(Fq* data, uint64_t n, const uint32_t* active) {
    if (!*active) return;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) data[i] = data[i].inv();
}
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
