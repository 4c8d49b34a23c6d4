// Host emulation of the experimental batched-affine bucket accumulation: the per-thread bodies of
// scroll-prover_b200/csrc/msm_affine.cuh run thread by thread (the threads of a level never communicate), orchestrated like
// msm_affine_accumulate in msm_affine.cu; the reference sums every bucket with the XYZZ mixed addition the default path uses.
#include "../../scroll-prover_b200/csrc/msm_affine.cuh"
#include <cstring>
#include <vector>
using namespace b200zk;

// out: NB affine points (identity = (0,0)); returns the number of levels run
extern "C" int msm_affine_host(const Affine* bases, const uint32_t* entries, const uint32_t* offsets, uint64_t NB, uint64_t max_entries,
                               uint32_t L, Affine* out) {
    uint32_t levels = 0;
    while ((1ull << levels) < max_entries) ++levels;
    const uint64_t cap1 = max_entries / 2 + NB + 1;
    std::vector<Affine> pts[2] = {std::vector<Affine>(cap1), std::vector<Affine>(cap1)};
    std::vector<uint32_t> offs[2] = {std::vector<uint32_t>(NB + 1), std::vector<uint32_t>(NB + 1)};
    std::vector<Fq> prefix(cap1), totals((cap1 + L - 1) / L + 1);
    BaLevel lv{bases, entries, nullptr, offsets, nullptr, NB};
    uint64_t bound = max_entries;
    for (uint32_t l = 0; l < levels; ++l) {
        std::vector<uint32_t>& oo = offs[l & 1];
        uint32_t acc = 0;
        for (uint64_t b = 0; b < NB; ++b) {  // what ba_next_counts_tiles / ba_scan_tiles / ba_next_offsets compute
            oo[b] = acc;
            acc += (lv.off_in[b + 1] - lv.off_in[b] + 1) >> 1;
        }
        oo[NB] = acc;
        lv.off_out = oo.data();
        uint64_t out_bound = bound / 2 + NB + 1;
        if (out_bound > cap1) out_bound = cap1;
        uint64_t nthreads = (out_bound + L - 1) / L;
        const bool active = acc < lv.off_in[NB];  // ba_scan_tiles' flag: inactive levels skip pass A and the inversion
        if (active) {
            for (uint64_t t = 0; t < nthreads; ++t) ba_thread_a(t, L, lv, prefix.data(), totals.data());
            for (uint64_t t = 0; t < nthreads; ++t) totals[t] = totals[t].inv();  // ba_invert_totals
        } else {
            for (auto& x : totals) x = Fq::zero();  // whatever is left there must not matter to pass B
            for (auto& x : prefix) x = Fq::zero();
        }
        for (uint64_t t = 0; t < nthreads; ++t) ba_thread_b(t, L, lv, prefix.data(), totals.data(), pts[l & 1].data());
        lv.entries = nullptr;
        lv.points = pts[l & 1].data();
        lv.off_in = oo.data();
        bound = out_bound;
    }
    for (uint64_t b = 0; b < NB; ++b) out[b] = xyzz_to_affine(ba_final_bucket(lv, b));
    return (int)levels;
}

extern "C" void msm_buckets_reference(const Affine* bases, const uint32_t* entries, const uint32_t* offsets, uint64_t NB, Affine* out) {
    for (uint64_t b = 0; b < NB; ++b) {
        XYZZ acc = XYZZ::identity();
        for (uint32_t i = offsets[b]; i < offsets[b + 1]; ++i) {
            Affine p = bases[entries[i] & 0x7fffffffu];
            if (p.is_identity()) continue;
            if (entries[i] & 0x80000000u) p.y = p.y.neg();
            xyzz_madd(acc, p.x, p.y);
        }
        out[b] = xyzz_to_affine(acc);
    }
}
