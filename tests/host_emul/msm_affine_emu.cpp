// The kernels of scroll-prover_b200/csrc/msm_affine_kernels.cuh compiled for the CPU under cuda_emu.hpp and driven exactly
// like msm_affine_accumulate (msm_affine.cu) drives them: per level the three offset-scan kernels (real thread blocks with
// barriers), the activity flag, pass A, the hierarchical inversion of the thread totals, pass B; then finalisation.
// tests/test_msm_affine_host.py compares the resulting buckets with the XYZZ reference sums.
#include "cuda_emu.hpp"
// clang-format off
#include "../../scroll-prover_b200/csrc/msm_affine_kernels.cuh"
// clang-format on
#include <vector>
using namespace b200zk;

static void emu_invert_totals(Fq* data, uint64_t n, Fq* scratch, const uint32_t* active) {
    if (n <= BAI_LEAF) {
        cuda_emu::launch((unsigned)((n + 127) / 128), 128, false, [&]() { bai_leaf(data, n, active); });
        return;
    }
    uint64_t T = (n + BAI_SLICE - 1) / BAI_SLICE;
    Fq *prefix = scratch, *totals = scratch + n;
    unsigned blocks = (unsigned)((T + 255) / 256);
    cuda_emu::launch(blocks, 256, false, [&]() { bai_up(data, prefix, totals, n, (uint32_t)T, active); });
    emu_invert_totals(totals, T, totals + T, active);
    cuda_emu::launch(blocks, 256, false, [&]() { bai_down(data, prefix, totals, n, (uint32_t)T, active); });
}
static size_t emu_invert_scratch(uint64_t n) {
    size_t e = 0;
    for (uint64_t m = n; m > BAI_LEAF;) {
        uint64_t T = (m + BAI_SLICE - 1) / BAI_SLICE;
        e += m + T;
        m = T;
    }
    return e + 1;
}

// mirrors msm_affine_accumulate; out: NB XYZZ buckets converted to affine for comparison; returns the number of ACTIVE levels
extern "C" int msm_affine_emu(const Affine* bases, const uint32_t* entries, const uint32_t* offsets, uint64_t NB, uint64_t max_entries,
                              uint32_t L, Affine* out) {
    const uint64_t cap1 = max_entries / 2 + NB + 1, tcap = (cap1 + L - 1) / L;
    const unsigned ntiles = (unsigned)((NB + BA_SCAN_TILE - 1) / BA_SCAN_TILE);
    std::vector<Affine> pts[2] = {std::vector<Affine>(cap1), std::vector<Affine>(cap1)};
    std::vector<uint32_t> offs[2] = {std::vector<uint32_t>(NB + 1), std::vector<uint32_t>(NB + 1)};
    std::vector<Fq> prefix(cap1), totals(tcap + 1), inv_scratch(emu_invert_scratch(tcap));
    std::vector<uint32_t> tiles(ntiles + 1);
    uint32_t active = 0;
    uint32_t levels = 0;
    while ((1ull << levels) < max_entries) ++levels;
    BaLevel lv{bases, entries, nullptr, offsets, nullptr, NB};
    uint64_t bound = max_entries;
    int active_levels = 0;
    for (uint32_t l = 0; l < levels; ++l) {
        uint32_t* off_out = offs[l & 1].data();
        Affine* o = pts[l & 1].data();
        cuda_emu::launch(ntiles, BA_SCAN_TPB, true, [&]() { ba_next_counts_tiles(lv.off_in, NB, tiles.data()); });
        cuda_emu::launch(1, 32, false, [&]() { ba_scan_tiles(tiles.data(), ntiles, off_out + NB, lv.off_in + NB, &active); });
        cuda_emu::launch(ntiles, BA_SCAN_TPB, true, [&]() { ba_next_offsets(lv.off_in, NB, tiles.data(), off_out); });
        lv.off_out = off_out;
        uint64_t out_bound = bound / 2 + NB + 1;
        if (out_bound > cap1) out_bound = cap1;
        uint64_t nthreads = (out_bound + L - 1) / L;
        unsigned blocks = (unsigned)((nthreads + 127) / 128);
        if (!active) {  // stale values from earlier levels stay in place on the device as well; poison them here
            for (auto& x : totals) x = Fq::zero();
            for (auto& x : prefix) x = Fq::zero();
        }
        active_levels += active ? 1 : 0;
        cuda_emu::launch(blocks, 128, false, [&]() { ba_pass_a(lv, prefix.data(), totals.data(), nthreads, L, &active); });
        emu_invert_totals(totals.data(), nthreads, inv_scratch.data(), &active);
        cuda_emu::launch(blocks, 128, false, [&]() { ba_pass_b(lv, prefix.data(), totals.data(), o, nthreads, L); });
        lv.entries = nullptr;
        lv.points = o;
        lv.off_in = off_out;
        bound = out_bound;
    }
    std::vector<XYZZ> buckets(NB);
    cuda_emu::launch((unsigned)((NB + 255) / 256), 256, false, [&]() { ba_finalize(lv, buckets.data()); });
    for (uint64_t b = 0; b < NB; ++b) out[b] = xyzz_to_affine(buckets[b]);
    return active_levels;
}
