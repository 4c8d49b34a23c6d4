// Kernels of the experimental batched-affine bucket accumulation (see msm_affine.cuh for the algorithm, msm_affine.cu for the
// host orchestration).  Kept in a header so that tests/host_emul/msm_affine_emu.cpp can compile the very same kernel source for
// the CPU under a small CUDA emulation layer (blocks of real threads, __syncthreads as a barrier) and run the whole level
// pipeline -- scans, flags, prefix passes, hierarchical inversion, finalisation -- without a GPU.
#pragma once
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
}  // namespace b200zk
