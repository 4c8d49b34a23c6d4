// BN254 G1 multi-scalar multiplication (Pippenger bucket method) for sm_100a.
//
// Device replacement for halo2_proofs::arithmetic::best_multiexp / multiexp_serial and therefore
// ParamsKZG::commit / commit_lagrange (halo2_proofs/src/arithmetic.rs, src/poly/kzg/commitment.rs @
// scroll-tech/halo2 e5ddf67, pin /root/reference/Cargo.lock:1886-1888; reached from
// /root/reference/integration/src/prove.rs:37-39).  The result is the same group element
// sum_i s_i * P_i, returned normalised, so its bytes equal the reference's `to_affine()` output.
//
// Pipeline (DESIGN.md "MSM"); all of it on the context stream, no host synchronisation inside:
//   1 msm_count       scalar -> canonical (one Montgomery product), signed c-bit digits -> digits[w*n+i], histogram
//   2 scan            exclusive prefix sum of the bucket counts (three kernels)
//   3 msm_scatter     counting sort of (table index | sign) into bucket order; window-major / a few bucket-range
//                     sweeps so that the random 4-byte stores stay L2-resident
//   4 msm_accumulate  lock-step segmented reduction: every thread owns L consecutive sorted entries and adds affine
//                     bases into an XYZZ accumulator (8M+2S per add); buckets wholly inside a chunk are stored
//                     directly, the <= 2 cut runs become partial records
//   5 msm_combine_*   partial records: ordinary buckets by their head record, GIANT buckets (> 4L entries: millions
//                     of equal digits in real witness columns) by a log-depth level reduction
//   6 msm_rowcol_sums + msm_bit_sums   sum_b b*B_b per bucket set with a short critical path (row / column sums, then
//                     bit-sliced weighted sums: independent subset sums instead of a dependent suffix scan)
//   7 msm_finish      one block per column: Horner over the bit sums and the bucket sets, normalisation to (x, y, 1)
// With a precomputed SRS (tables 2^(c*w) P_i, built at registration) all windows share ONE bucket set.
// BATCHES: up to 32 columns over the same bases run through one pipeline (bucket set = column * Ws + window; the grids of
// count / scatter carry the column in blockIdx.y), so the fixed phases are paid once per batch (msm_run_batch).
// Zero scalars and zero digits are skipped (witness columns are mostly zeros / small values).
#include "common.cuh"
#include "ec.cuh"

namespace b200zk {

static constexpr uint32_t PART_INVALID = 0x1fffffffu;  // also the bucket-id mask
static constexpr uint32_t PART_GIANT = 0x20000000u;
static constexpr uint32_t PART_STARTS = 0x80000000u;
static constexpr uint32_t PART_ENDS = 0x40000000u;
static constexpr int MSM_MAX_BATCH = 32;
static constexpr int ACC_L_DEFAULT = 256;  // sorted entries per accumulate thread (B200ZK_ACC_L overrides, experiments)

struct MsmPlan {
    uint32_t c, W, B;   // window bits, windows, buckets per window (2^(c-1))
    uint32_t Ws;        // bucket sets per column: W, or 1 when the SRS holds precomputed 2^(c*w) multiples of every base
    uint32_t batch;     // columns (scalar vectors over the SAME bases) summed in one pipeline; bucket set = col * Ws + w
    uint64_t stride;    // precomputed SRS: table w starts at bases + w*stride (0 otherwise)
    uint64_t NB;        // batch * Ws * B
};

// Batched MSM: up to MSM_MAX_BATCH columns over the same bases go through ONE count / sort / accumulate / reduce pipeline
// (their bucket sets lie side by side), so the latency-bound phases -- scans, the bucket reduction, the final Horner --
// are paid once per batch instead of once per column.  This is what the hundreds of 2^20-row columns of the inner
// (zkEVM super-circuit) proof need; a 2^24+ column fills the machine alone and runs with batch = 1.
struct MsmCols {
    const Fr* p[MSM_MAX_BATCH];
};

__device__ __forceinline__ void ld_affine(const Affine* p, Fq& x, Fq& y) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2), d = __ldg(q + 3);
    x.l.v[0] = a.x; x.l.v[1] = a.y; x.l.v[2] = a.z; x.l.v[3] = a.w;
    x.l.v[4] = b.x; x.l.v[5] = b.y; x.l.v[6] = b.z; x.l.v[7] = b.w;
    y.l.v[0] = c.x; y.l.v[1] = c.y; y.l.v[2] = c.z; y.l.v[3] = c.w;
    y.l.v[4] = d.x; y.l.v[5] = d.y; y.l.v[6] = d.z; y.l.v[7] = d.w;
}

// every scalar once: canonical form, signed digits -> digits[w*n + i] (magnitude | sign<<31, 0 = skip) + histogram
__global__ void __launch_bounds__(256) msm_count(MsmCols cols, uint64_t n, MsmPlan pl, uint32_t* hist, uint32_t* digits_all) {
    // grid = (blocks over the scalars, column of the batch): no 64-bit division per element
    const uint32_t col = blockIdx.y;
    const Fr* scalars = cols.p[0];
#pragma unroll
    for (int q = 1; q < MSM_MAX_BATCH; ++q)
        if (q == (int)col) scalars = cols.p[q];  // no dynamic indexing of kernel parameters
    uint32_t* digits = digits_all + (uint64_t)col * pl.W * n;
    uint32_t* hist_c = hist + (uint64_t)col * pl.Ws * pl.B;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr s = scalars[i];
        if (s.is_zero()) {
            for (uint32_t w = 0; w < pl.W; ++w) digits[(uint64_t)w * n + i] = 0;
            continue;
        }
        s = s.from_mont();
        uint32_t l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = s.l.v[k];
        const uint32_t c = pl.c, mask = (1u << c) - 1, half = 1u << (c - 1);
        uint32_t carry = 0;
        for (uint32_t w = 0; w < pl.W; ++w) {
            uint32_t v = (l[0] & mask) + carry;
#pragma unroll
            for (int k = 0; k < 7; ++k) l[k] = __funnelshift_r(l[k], l[k + 1], c);
            l[7] >>= c;
            uint32_t enc;
            if (v > half) {
                enc = (1u << c) - v;  // magnitude of the negative digit (0 when v == 2^c)
                carry = 1;
                if (enc) enc |= 0x80000000u;
            } else {
                enc = v;
                carry = 0;
            }
            digits[(uint64_t)w * n + i] = enc;
            if (enc) atomicAdd(&hist_c[(pl.Ws == 1 ? 0ull : (uint64_t)w * pl.B) + (enc & 0x7fffffffu) - 1], 1u);
        }
    }
}

// counting-sort scatter in WINDOW-MAJOR order: concurrently running blocks work on the same window, so the
// random 4-byte stores fall into one n*4-byte region that stays L2-resident (64 MiB at n = 2^24)
// With a precomputed SRS (one bucket set) the destination of a digit is spread over the whole n*W*4-byte entry
// array, so the kernel is run in SWEEPS over bucket ranges [b_lo, b_hi): each sweep re-reads the digits (coalesced)
// but writes into a region small enough to stay L2-resident until its sectors are complete.
__global__ void __launch_bounds__(256) msm_scatter(const uint32_t* __restrict__ digits, uint64_t n, MsmPlan pl,
                                                   uint32_t* cursor, uint32_t* entries, const uint32_t* __restrict__ total_entries,
                                                   uint32_t sweep, uint32_t max_sweeps) {
    // the number of sweeps actually used follows the real entry count M (known only on the device): sparse
    // witness columns have few entries and get a single sweep
    uint32_t eff = 1;
    if (max_sweeps > 1) {
        uint64_t bytes = 4ull * (*total_entries);
        eff = (uint32_t)((bytes + (1ull << 28) - 1) >> 28);  // regions of <= 256 MiB (a shift: this runs once per thread)
        eff = eff < 1 ? 1 : (eff > max_sweeps ? max_sweeps : eff);
    }
    if (sweep >= eff) return;
    // B is a power of two and eff <= 4: the range bounds need no 64-bit division either
    const uint32_t b_lo = eff == 3 ? (pl.B / 3) * sweep : (pl.B / eff) * sweep;
    const uint32_t b_hi = (sweep + 1 == eff) ? pl.B : (eff == 3 ? (pl.B / 3) * (sweep + 1) : (pl.B / eff) * (sweep + 1));
    // grid = (blocks over the scalars, col * W + w): the digits of one (column, window) are a contiguous run, so neither the
    // window nor the column needs a 64-bit division per digit; blocks are still issued window-major (x fastest)
    const uint32_t cw = blockIdx.y;
    const uint32_t col = cw / pl.W, w = cw - col * pl.W;
    const uint32_t* dg = digits + (uint64_t)cw * n;
    uint32_t* cur = cursor + ((uint64_t)col * pl.Ws + (pl.Ws == 1 ? 0u : w)) * pl.B;
    const uint32_t woff = (uint32_t)((uint64_t)w * pl.stride);
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint32_t d = dg[i];
        if (!d) continue;
        uint32_t bk = (d & 0x7fffffffu) - 1;
        if (bk < b_lo || bk >= b_hi) continue;
        uint32_t pos = atomicAdd(&cur[bk], 1u);
        entries[pos] = ((uint32_t)i + woff) | (d & 0x80000000u);
    }
}

// ---- exclusive scan of uint32 counts (three kernels) -----------------------------------------
static constexpr int SCAN_TPB = 256, SCAN_ITEMS = 16, SCAN_TILE = SCAN_TPB * SCAN_ITEMS;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t warp_sums[SCAN_TPB / 32];
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = (lane < SCAN_TPB / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= (uint32_t)o) s += y;
        }
        if (lane < SCAN_TPB / 32) warp_sums[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    uint32_t warp_off = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_TPB / 32 - 1];
    uint32_t r = warp_off + x - v;
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_TPB) scan_tile_sums(const uint32_t* in, uint64_t n, uint32_t* tile_sums) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    uint32_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}
// one block: exclusive scan of the tile sums in chunks of SCAN_TPB with a running carry
__global__ void __launch_bounds__(SCAN_TPB) scan_tile_offsets(uint32_t* tile_sums, uint32_t ntiles, uint32_t* grand_total,
                                                              unsigned long long* running) {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < ntiles; base += SCAN_TPB) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0, total;
        uint32_t ex = block_exclusive_scan(v, &total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        *grand_total = carry;
        if (running) atomicAdd(running, (unsigned long long)carry);
    }
}
__global__ void __launch_bounds__(SCAN_TPB) scan_apply(const uint32_t* in, uint64_t n, const uint32_t* tile_offs, uint32_t* out,
                                                     uint32_t* out2) {
    uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE + (uint64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t off = block_exclusive_scan(s, &total) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) {
            out[base + k] = off;
            out2[base + k] = off;
        }
        off += v[k];
    }
}

// ---- bucket accumulation ------------------------------------------------------------------------
__device__ __forceinline__ void st_xyzz(XYZZ* p, const XYZZ& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    const Fq* f = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        q[2 * k] = make_uint4(f[k].l.v[0], f[k].l.v[1], f[k].l.v[2], f[k].l.v[3]);
        q[2 * k + 1] = make_uint4(f[k].l.v[4], f[k].l.v[5], f[k].l.v[6], f[k].l.v[7]);
    }
}
__device__ __forceinline__ XYZZ ld_xyzz(const XYZZ* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    XYZZ v;
    Fq* f = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint4 a = q[2 * k], b = q[2 * k + 1];
        f[k].l.v[0] = a.x; f[k].l.v[1] = a.y; f[k].l.v[2] = a.z; f[k].l.v[3] = a.w;
        f[k].l.v[4] = b.x; f[k].l.v[5] = b.y; f[k].l.v[6] = b.z; f[k].l.v[7] = b.w;
    }
    return v;
}

// offsets has NB + 1 entries (offsets[NB] = M sorted entries).  Perfect load balance: thread tau sums exactly the
// sorted entries [tau*L, (tau+1)*L) in lock step with its warp (one mixed add per iteration for every lane).
// Buckets that lie wholly inside the range are stored directly; the (at most two) runs cut by the range ends are
// emitted as partial records.  Records of ordinary buckets are merged by msm_combine_heads (a bucket of <= BIG
// entries is cut into <= BIG/L + 1 pieces); records of GIANT buckets (millions of equal digits in real witness
// columns) go through the log-depth msm_combine_level reduction instead.
__global__ void __launch_bounds__(256, 2)
msm_accumulate(const Affine* __restrict__ bases, const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets,
               uint64_t NB, XYZZ* __restrict__ buckets, uint32_t* __restrict__ part_id, XYZZ* __restrict__ part_val,
               uint64_t nthreads, uint32_t* __restrict__ giant_flag, uint32_t ACC_L) {
    uint64_t tau = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tau >= nthreads) return;
    const uint32_t M = offsets[NB];
    const uint32_t BIG_BUCKET = 4 * ACC_L;  // buckets above this size are split over threads
    uint64_t start = tau * ACC_L;
    part_id[2 * tau] = PART_INVALID;
    part_id[2 * tau + 1] = PART_INVALID;
    if (start >= M) return;
    uint32_t end = (start + ACC_L < M) ? (uint32_t)(start + ACC_L) : M;
    uint64_t lo = 0, hi = NB;  // largest b with offsets[b] <= start
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t b = (uint32_t)lo;
    uint32_t bucket_start = offsets[b], bucket_end = offsets[b + 1];
    bool run_starts = (bucket_start == (uint32_t)start);
    uint32_t slot = 0;
    XYZZ acc = XYZZ::identity();
    uint32_t pos = (uint32_t)start;
    while (true) {
        bool at_end = (pos == end);
        if (at_end || pos == bucket_end) {
            bool run_ends = (pos == bucket_end);
            if (run_starts && run_ends) {
                st_xyzz(buckets + b, acc);
            } else {
                uint32_t giant = (bucket_end - bucket_start > BIG_BUCKET) ? PART_GIANT : 0u;
                if (giant) *giant_flag = 1u;  // benign race: every writer stores the same value
                part_id[2 * tau + slot] = b | giant | (run_starts ? PART_STARTS : 0u) | (run_ends ? PART_ENDS : 0u);
                st_xyzz(part_val + 2 * tau + slot, acc);
                ++slot;
            }
            if (at_end) break;
            acc = XYZZ::identity();
            bucket_start = pos;
            do { ++b; bucket_end = offsets[b + 1]; } while (bucket_end == pos);  // skip empty buckets
            run_starts = true;
        }
        uint32_t e = entries[pos];
        Fq px, py;
        ld_affine(bases + (e & 0x7fffffffu), px, py);
        if (!(px.is_zero() && py.is_zero())) {
            if (e & 0x80000000u) py = py.neg();
            xyzz_madd(acc, px, py);
        }
        ++pos;
    }
}

// ordinary (non-giant) buckets: the head record sums the few following records up to the one that ends the bucket
__global__ void __launch_bounds__(128) msm_combine_heads(const uint32_t* __restrict__ part_id, const XYZZ* __restrict__ part_val,
                                                         uint64_t nrec, XYZZ* __restrict__ buckets) {
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    uint32_t id = part_id[r];
    if ((id & PART_INVALID) == PART_INVALID || (id & PART_GIANT) || !(id & PART_STARTS)) return;
    XYZZ sum = ld_xyzz(part_val + r);
    uint64_t k = r;
    while (!(id & PART_ENDS)) {
        ++k;
        if (k >= nrec) break;
        id = part_id[k];
        if ((id & PART_INVALID) == PART_INVALID) { id = 0; continue; }
        XYZZ v = ld_xyzz(part_val + k);
        xyzz_add(sum, v);
    }
    st_xyzz(buckets + (part_id[r] & PART_INVALID), sum);
}

// One level of the GIANT-bucket reduction: thread sigma scans LR consecutive records (sorted by bucket; invalid and,
// on the first level, non-giant records are skipped), sums runs of equal bucket id; a run that saw both the STARTS
// and the ENDS record is complete and is stored, otherwise it is re-emitted (<= 2 per thread) for the next level.
static constexpr int COMBINE_LR = 64;
__global__ void __launch_bounds__(128) msm_combine_level(const uint32_t* __restrict__ in_id, const XYZZ* __restrict__ in_val,
                                                         uint64_t nrec, XYZZ* __restrict__ buckets, uint32_t* __restrict__ out_id,
                                                         XYZZ* __restrict__ out_val, uint64_t nthreads,
                                                         const uint32_t* __restrict__ giant_flag) {
    if (*giant_flag == 0) return;  // no giant bucket in this MSM (the common case): nothing to reduce
    uint64_t sigma = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (sigma >= nthreads) return;
    out_id[2 * sigma] = PART_INVALID;
    out_id[2 * sigma + 1] = PART_INVALID;
    uint64_t r0 = sigma * COMBINE_LR, r1 = r0 + COMBINE_LR;
    if (r1 > nrec) r1 = nrec;
    bool have = false;
    uint32_t cur = 0, flags = 0, slot = 0;
    XYZZ acc = XYZZ::identity();
    auto flush = [&]() {
        if (!have) return;
        if ((flags & PART_STARTS) && (flags & PART_ENDS)) {
            st_xyzz(buckets + cur, acc);
        } else {
            out_id[2 * sigma + slot] = cur | flags | PART_GIANT;
            st_xyzz(out_val + 2 * sigma + slot, acc);
            ++slot;
        }
    };
    for (uint64_t r = r0; r < r1; ++r) {
        uint32_t id = in_id[r];
        if ((id & PART_INVALID) == PART_INVALID || !(id & PART_GIANT)) continue;
        uint32_t bkt = id & PART_INVALID;
        if (!have || bkt != cur) {
            flush();
            have = true;
            cur = bkt;
            flags = id & (PART_STARTS | PART_ENDS);
            acc = ld_xyzz(in_val + r);
        } else {
            flags |= id & (PART_STARTS | PART_ENDS);
            XYZZ v = ld_xyzz(in_val + r);
            xyzz_add(acc, v);
        }
    }
    flush();
}

// last level (few records): the head record of every giant bucket sums forward to the record that ends it
__global__ void __launch_bounds__(128) msm_combine_final(const uint32_t* __restrict__ part_id, const XYZZ* __restrict__ part_val,
                                                         uint64_t nrec, XYZZ* __restrict__ buckets, const uint32_t* __restrict__ giant_flag) {
    if (*giant_flag == 0) return;
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    uint32_t id = part_id[r];
    if ((id & PART_INVALID) == PART_INVALID || !(id & PART_GIANT) || !(id & PART_STARTS)) return;
    XYZZ sum = ld_xyzz(part_val + r);
    uint64_t k = r;
    while (!(id & PART_ENDS)) {
        ++k;
        if (k >= nrec) break;
        id = part_id[k];
        if ((id & PART_INVALID) == PART_INVALID || !(id & PART_GIANT)) { id = 0; continue; }
        XYZZ v = ld_xyzz(part_val + k);
        xyzz_add(sum, v);
    }
    st_xyzz(buckets + (part_id[r] & PART_INVALID), sum);
}

// ---- bucket reduction ---------------------------------------------------------------------------
// S = sum_{i<B} (i+1) B_i for every bucket set, with a SHORT critical path (the old running-sum + doubling tree had
// ~100 dependent point additions and ~200 dependent doublings; this has ~40 additions and kc doublings):
// view the set as a matrix i = hi*2^kc + lo;  Row_hi = sum_lo B, Col_lo = sum_hi B   (block tree sums), then
//   S = WS(Col) + 2^kc * WS(Row) + sum(Row),     WS(V) = sum_j j V_j = sum_{j>=1} Suffix_j(V)   (parallel suffix scan).
static constexpr int RED_T = 128;   // 4 blocks/SM at 128 registers: the sums are latency-bound, more blocks in flight win

// block tree sum through shared memory (result valid in thread 0).  A register-shuffle version (32 x SHFL.DOWN per level and
// lane, one shared-memory hand-over between the warps) was measured on the B200 and is SLOWER: bucket reduction of 2^21 buckets
// 2.92 ms against 2.10 ms with this one -- every lane of a shuffle level executes the 14-multiplication addition (31 of 32
// results are discarded at the last level) and the extra live XYZZ pushed the kernels into spills at 128 registers.
__device__ __forceinline__ void block_tree_sum(XYZZ& v, XYZZ* sh) {  // result valid in thread 0
    st_xyzz(sh + threadIdx.x, v);
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            XYZZ a = ld_xyzz(sh + threadIdx.x), b = ld_xyzz(sh + threadIdx.x + s);
            xyzz_add(a, b);
            st_xyzz(sh + threadIdx.x, a);
        }
        __syncthreads();
    }
    v = ld_xyzz(sh);
}

// grid = (rows + cols, Ws); block j < rows sums row j, block rows + j sums column j.  vec[set][0..rows) | [rows..rows+cols)
__global__ void __launch_bounds__(RED_T, 4) msm_rowcol_sums(const XYZZ* __restrict__ buckets, uint32_t B, uint32_t kc, XYZZ* __restrict__ vec) {
    __shared__ XYZZ sh[RED_T];
    const uint32_t cols = 1u << kc, rows = B >> kc;
    const XYZZ* bk = buckets + (uint64_t)blockIdx.y * B;
    XYZZ acc = XYZZ::identity();
    if (blockIdx.x < rows) {
        const XYZZ* r = bk + (uint64_t)blockIdx.x * cols;
        for (uint32_t lo = threadIdx.x; lo < cols; lo += blockDim.x) {
            XYZZ v = ld_xyzz(r + lo);
            xyzz_add(acc, v);
        }
    } else {
        uint32_t lo = blockIdx.x - rows;
        for (uint32_t hi = threadIdx.x; hi < rows; hi += blockDim.x) {
            XYZZ v = ld_xyzz(bk + (uint64_t)hi * cols + lo);
            xyzz_add(acc, v);
        }
    }
    block_tree_sum(acc, sh);
    if (threadIdx.x == 0) st_xyzz(vec + (uint64_t)blockIdx.y * (rows + cols) + blockIdx.x, acc);
}

// WS(V) = sum_j j V_j over a vector of m = 2^q elements, BIT-SLICED:  WS = sum_{b<q} 2^b S_b,  S_b = sum_{j: bit b of j} V_j.
// The q subset sums are independent block tree sums (depth ~ m/(2*RED_T) + log2 RED_T additions instead of the ~90 dependent
// additions of a suffix scan); msm_finish folds them with q - 1 doublings.
// grid = (q_max + 1, 2, sets): blockIdx.y = 0 the Row vector (m = rows), 1 the Col vector (m = cols); blockIdx.x = bit b, and
// blockIdx.x == q_max of the Row vector computes its plain total sum(Row).  out[set][which][b], stride (q_max + 1).
__global__ void __launch_bounds__(RED_T, 4) msm_bit_sums(const XYZZ* __restrict__ vec, uint32_t B, uint32_t kc, uint32_t q_max,
                                                         XYZZ* __restrict__ out) {
    __shared__ XYZZ sh[RED_T];
    const uint32_t cols = 1u << kc, rows = B >> kc;
    const uint32_t which = blockIdx.y, b = blockIdx.x;
    const uint32_t m = which == 0 ? rows : cols;
    const XYZZ* v = vec + (uint64_t)blockIdx.z * (rows + cols) + (which == 0 ? 0 : rows);
    XYZZ acc = XYZZ::identity();
    const bool total = (b == q_max);
    if (total ? (which == 0) : ((1u << b) < m)) {
        for (uint32_t j = threadIdx.x; j < m; j += blockDim.x)
            if (total || ((j >> b) & 1u)) {
                XYZZ e = ld_xyzz(v + j);
                xyzz_add(acc, e);
            }
    }
    block_tree_sum(acc, sh);
    if (threadIdx.x == 0) st_xyzz(out + ((uint64_t)blockIdx.z * 2 + which) * (q_max + 1) + b, acc);
}

// one block per column of the batch: warp 0 folds the Row bit sums, warp 1 the Col bit sums (Horner with doublings), then
// thread 0 combines  S = WS(Col) + 2^kc WS(Row) + sum(Row)  per bucket set, the window Horner, and normalises.
__global__ void __launch_bounds__(64) msm_finish(MsmPlan pl, uint32_t kc, uint32_t q_max, const XYZZ* __restrict__ bits_all, Jacobian* out_all) {
    __shared__ XYZZ ws[2];
    const uint32_t cols = 1u << kc, rows = pl.B >> kc;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Jacobian* out = out_all + blockIdx.x;
    XYZZ acc = XYZZ::identity();
    for (uint32_t w = pl.Ws; w-- > 0;) {
        const uint64_t set = (uint64_t)blockIdx.x * pl.Ws + w;
        const XYZZ* bits = bits_all + set * 2 * (q_max + 1);
        if (lane == 0) {
            const uint32_t m = warp == 0 ? rows : cols;
            const XYZZ* bv = bits + (uint64_t)warp * (q_max + 1);
            XYZZ h = XYZZ::identity();
            uint32_t q = 0;
            while ((1u << q) < m) ++q;
            for (uint32_t b = q; b-- > 0;) {
                if (!h.is_identity()) h = xyzz_dbl(h);
                XYZZ e = ld_xyzz(bv + b);
                xyzz_add(h, e);
            }
            st_xyzz(ws + warp, h);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (!acc.is_identity())                                    // (always the identity when Ws == 1: skip the c doublings)
                for (uint32_t d = 0; d < pl.c; ++d) acc = xyzz_dbl(acc);
            XYZZ s = ld_xyzz(ws);                                      // WS(Row)
            for (uint32_t d = 0; d < kc; ++d) s = xyzz_dbl(s);
            XYZZ t = ld_xyzz(bits + q_max), u = ld_xyzz(ws + 1);       // sum(Row), WS(Col)
            xyzz_add(s, t);
            xyzz_add(s, u);
            xyzz_add(acc, s);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = xyzz_to_jacobian_normalized(acc);
}

// ---- small helpers exposed through the ABI --------------------------------------------------------
__global__ void g1_sum_kernel(const Jacobian* pts, uint64_t count, Jacobian* out) {
    if (threadIdx.x || blockIdx.x) return;
    XYZZ acc = XYZZ::identity();
    for (uint64_t i = 0; i < count; ++i) {
        XYZZ p = xyzz_from_jacobian(pts[i]);
        xyzz_add(acc, p);
    }
    *out = xyzz_to_jacobian_normalized(acc);
}

__global__ void __launch_bounds__(128) g1_generator_mul_kernel(const Fr* scalars, uint64_t n, Affine* out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr s = scalars[i].from_mont();
    Fq gx = Fq::one(), gy = Fq::one().dbl();  // generator (1, 2)
    XYZZ acc = XYZZ::identity();
    for (int limb = 7; limb >= 0; --limb)
        for (int b = 31; b >= 0; --b) {
            acc = xyzz_dbl(acc);
            if ((s.l.v[limb] >> b) & 1) xyzz_madd(acc, gx, gy);
        }
    out[i] = xyzz_to_affine(acc);
}

// ---------------------------------------------------------------------------------------------
static uint32_t pick_window(uint64_t n) {
    if (n < 32) return n < 4 ? 2 : 4;
    double best = 1e300;
    uint32_t bc = 8;
    for (uint32_t c = 5; c <= 23; ++c) {
        uint32_t W = 254 / c + 1;
        if ((double)n * W >= 4.0e9) continue;  // sorted-entry positions are 32-bit
        double cost = (double)n * W * 10.0 + (double)W * (double)(1ull << c) * 40.0 + (double)W * 4096.0;
        if (cost < best) {
            best = cost;
            bc = c;
        }
    }
    return bc;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// out[i] = 2^c * in[i] (affine): one table of the precomputed SRS from the previous one
__global__ void __launch_bounds__(128) srs_shift_kernel(const Affine* __restrict__ in, Affine* __restrict__ out, uint64_t n, uint32_t c) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine a = in[i];
    XYZZ p = xyzz_from_affine(a);
    for (uint32_t d = 0; d < c; ++d) p = xyzz_dbl(p);
    out[i] = xyzz_to_affine(p);
}

uint32_t msm_pick_window_precomputed(uint64_t n) {
    double best = 1e300;
    uint32_t bc = 10;
    for (uint32_t c = 8; c <= 23; ++c) {
        uint32_t W = 254 / c + 1;
        if ((double)n * W >= 2.0e9) continue;  // entry = table index (31 bits) | sign
        // 10 MODMUL per bucket addition; ~40 MODMUL-equivalents per bucket slot for the (latency-bound) reduction,
        // measured: 2.65 ms for 2^21 buckets next to 32 ms for 201 M additions
        double cost = (double)n * W * 10.0 + (double)(1ull << c) * 40.0;
        if (cost < best) {
            best = cost;
            bc = c;
        }
    }
    return bc;
}

// tables[w] = 2^(c*w) * bases, w = 0..W-1, laid out back to back (stride n); tables[0] must already hold the bases
int32_t srs_precompute_run(b200zk_ctx* ctx, Affine* tables, uint64_t n, uint32_t c, uint32_t W) {
    for (uint32_t w = 1; w < W; ++w) {
        srs_shift_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(tables + (uint64_t)(w - 1) * n, tables + (uint64_t)w * n, n, c);
        B2_LAUNCH_CHECK(ctx);
    }
    return B200ZK_OK;
}

// how many columns of n scalars one batched pipeline may take (sorted-entry positions are 32-bit; scratch stays ~2 GiB)
uint32_t msm_max_batch(uint64_t n, uint32_t pre_c) {
    uint32_t c = pre_c ? pre_c : pick_window(n);
    uint64_t per_col = (n ? n : 1) * (254 / c + 1);
    uint64_t b = (1ull << 28) / per_col;
    return (uint32_t)(b < 1 ? 1 : (b > MSM_MAX_BATCH ? MSM_MAX_BATCH : b));
}

// pre_c != 0: `bases` is a precomputed SRS (W tables of stride pre_stride) built for window pre_c.
// cols[0 .. batch): device pointers of `batch` scalar vectors of n elements each; out_dev[0 .. batch).
int32_t msm_run_batch(b200zk_ctx* ctx, const Affine* bases, const Fr* const* cols, uint32_t batch, uint64_t n, Jacobian* out_dev,
                      uint32_t pre_c, uint64_t pre_stride) {
    if (n >= (1ull << 31)) return fail(ctx, B200ZK_E_UNSUPPORTED, "msm: n = %llu >= 2^31", (unsigned long long)n);
    if (batch < 1 || batch > (uint32_t)MSM_MAX_BATCH) return fail(ctx, B200ZK_E_INVALID, "msm: batch %u out of range [1,%d]", batch, MSM_MAX_BATCH);
    MsmPlan pl;
    pl.c = pre_c ? pre_c : (ctx->msm_window ? ctx->msm_window : pick_window(n));
    if (pl.c < 2 || pl.c > 24) return fail(ctx, B200ZK_E_INVALID, "msm: window %u out of range [2,24]", pl.c);
    pl.W = 254 / pl.c + 1;
    pl.B = 1u << (pl.c - 1);
    pl.Ws = pre_c ? 1 : pl.W;
    pl.batch = batch;
    pl.stride = pre_c ? pre_stride : 0;
    pl.NB = (uint64_t)batch * pl.Ws * pl.B;
    uint64_t max_entries = n * pl.W * batch;
    if (max_entries >= 0xffffffffull) return fail(ctx, B200ZK_E_UNSUPPORTED, "msm: n*W*batch too large for window %u", pl.c);
    MsmCols colp;
    for (int q = 0; q < MSM_MAX_BATCH; ++q) colp.p[q] = cols[q < (int)batch ? q : 0];
    ctx->last_c = pl.c;
    ctx->last_windows = pl.W;
    ctx->last_adds = n * pl.W;

    const uint32_t ACC_L = ctx->msm_acc_l ? ctx->msm_acc_l : (uint32_t)ACC_L_DEFAULT;
    uint64_t nthreads = (max_entries + ACC_L - 1) / ACC_L;
    if (nthreads == 0) nthreads = 1;
    uint32_t ntiles = (uint32_t)((pl.NB + SCAN_TILE - 1) / SCAN_TILE);
    // carve the scratch arena
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_hist = carve(4 * (pl.NB + 1)), o_offs = carve(4 * (pl.NB + 1)), o_cursor = carve(4 * (pl.NB + 1));
    size_t o_tiles = carve(4 * (size_t)(ntiles + 1));
    size_t o_flag = carve(256);
    size_t o_entries = carve(4 * (max_entries + 4));
    size_t o_digits = carve(4 * (max_entries + 4));
    size_t o_buckets = carve(sizeof(XYZZ) * pl.NB);
    size_t o_pid = carve(4 * 2 * nthreads), o_pval = carve(sizeof(XYZZ) * 2 * nthreads);
    uint64_t nthreads2 = (2 * nthreads + COMBINE_LR - 1) / COMBINE_LR;
    size_t o_pid2 = carve(4 * 2 * nthreads2), o_pval2 = carve(sizeof(XYZZ) * 2 * nthreads2);
    const uint32_t kc = pl.c / 2;                       // columns = 2^kc, rows = B / 2^kc  (c - 1 = kc + kr)
    const uint32_t red_cols = 1u << kc, red_rows = pl.B >> kc;
    const uint32_t sets = batch * pl.Ws;
    size_t red_len = (size_t)sets * (red_rows + red_cols);
    uint32_t q_max = 0;  // bits of the longer of the Row / Col vectors
    while ((1u << q_max) < (red_rows > red_cols ? red_rows : red_cols)) ++q_max;
    size_t o_gr = carve(sizeof(XYZZ) * red_len), o_sums = carve(sizeof(XYZZ) * 2 * (q_max + 1) * sets);
    B2_TRY(scratch_reserve(ctx, ctx->msm_work, off));
    char* base = (char*)ctx->msm_work.p;
    uint32_t* hist = (uint32_t*)(base + o_hist);
    uint32_t* offsets = (uint32_t*)(base + o_offs);
    uint32_t* cursor = (uint32_t*)(base + o_cursor);
    uint32_t* tiles = (uint32_t*)(base + o_tiles);
    uint32_t* giant_flag = (uint32_t*)(base + o_flag);
    uint32_t* entries = (uint32_t*)(base + o_entries);
    uint32_t* digits = (uint32_t*)(base + o_digits);
    uint32_t* pid2 = (uint32_t*)(base + o_pid2);
    XYZZ* pval2 = (XYZZ*)(base + o_pval2);
    XYZZ* buckets = (XYZZ*)(base + o_buckets);
    uint32_t* pid = (uint32_t*)(base + o_pid);
    XYZZ* pval = (XYZZ*)(base + o_pval);
    XYZZ* grpR = (XYZZ*)(base + o_gr);
    XYZZ* red_sums = (XYZZ*)(base + o_sums);

    cudaStream_t st = ctx->stream;
    B2_CUDA(ctx, cudaMemsetAsync(hist, 0, 4 * (pl.NB + 1), st));
    B2_CUDA(ctx, cudaMemsetAsync(giant_flag, 0, 4, st));
    B2_CUDA(ctx, cudaMemsetAsync(buckets, 0, sizeof(XYZZ) * pl.NB, st));
    uint32_t sblocks = (uint32_t)ctx->sm_count * 8;
    if (n) {
        uint64_t want = (n + 255) / 256;
        uint32_t per_col = (sblocks + batch - 1) / batch;
        uint32_t blocks = (uint32_t)(want < per_col ? want : per_col);
        {
            ProfScope ps_(ctx, PROF_MSM_COUNT);
            msm_count<<<dim3(blocks, batch), 256, 0, st>>>(colp, n, pl, hist, digits);
        }
        B2_LAUNCH_CHECK(ctx);
    }
    {
        ProfScope ps_(ctx, PROF_MSM_SCAN);
        scan_tile_sums<<<ntiles, SCAN_TPB, 0, st>>>(hist, pl.NB, tiles);
        B2_LAUNCH_CHECK(ctx);
        scan_tile_offsets<<<1, SCAN_TPB, 0, st>>>(tiles, ntiles, offsets + pl.NB, ctx->msm_adds_dev);
        B2_LAUNCH_CHECK(ctx);
        scan_apply<<<ntiles, SCAN_TPB, 0, st>>>(hist, pl.NB, tiles, offsets, cursor);
        B2_LAUNCH_CHECK(ctx);
    }
    if (n) {
        {
            ProfScope ps_(ctx, PROF_MSM_SCATTER);
            uint64_t want = (n + 1023) / 1024;  // 4 digits per thread; grid.y = col * W + w, issued window-major (x fastest)
            uint32_t blocks = (uint32_t)(want < 0x7fffffffull ? (want ? want : 1) : 0x7fffffffull);
            {
                uint32_t sweeps = 1;
                if (pl.Ws == 1 && batch == 1) {
                    // measured on B200 at n = 2^24 (805 MB of entries): 1 sweep 6.3 ms, 2: 4.9, 4: 4.1, 16: 11.5 -- each
                    // sweep re-reads the digits, so only a few sweeps of ~200 MB pay off
                    uint64_t region = 200ull << 20;
                    sweeps = (uint32_t)((4 * max_entries + region - 1) / region);
                    if (sweeps < 1) sweeps = 1;
                    if (sweeps > 4) sweeps = 4;
                    if (ctx->msm_scatter_sweeps) sweeps = ctx->msm_scatter_sweeps;
                }
                for (uint32_t sw = 0; sw < sweeps; ++sw) {
                    // sweeps after the first may turn out to be unnecessary (the real entry count is only known on the device:
                    // a witness-like column needs one): they get a grid-stride launch of a few blocks per SM, so that an
                    // early exit costs microseconds instead of ~0.1 ms for ~200 K empty blocks (measured: witness-like 2^24
                    // scatter 1.16 -> 0.84 ms, whole MSM 9.06 -> 8.67 ms)
                    uint32_t bx = blocks;
                    if (sw > 0) {
                        uint32_t lean = (uint32_t)ctx->sm_count * 32u / (batch * pl.W) + 1;
                        if (lean < bx) bx = lean;
                    }
                    msm_scatter<<<dim3(bx, batch * pl.W), 256, 0, st>>>(digits, n, pl, cursor, entries, offsets + pl.NB, sw, sweeps);
                    if (sw + 1 < sweeps) B2_LAUNCH_CHECK(ctx);
                }
            }
        }
        B2_LAUNCH_CHECK(ctx);
        uint32_t ablocks = (uint32_t)((nthreads + 255) / 256);
        {
            ProfScope ps_(ctx, PROF_MSM_ACCUM);
            msm_accumulate<<<ablocks, 256, 0, st>>>(bases, entries, offsets, pl.NB, buckets, pid, pval, nthreads, giant_flag, ACC_L);
        }
        B2_LAUNCH_CHECK(ctx);
        {
            ProfScope ps_(ctx, PROF_MSM_COMBINE);
            uint64_t nrec = 2 * nthreads;
            msm_combine_heads<<<(uint32_t)((nrec + 127) / 128), 128, 0, st>>>(pid, pval, nrec, buckets);
            B2_LAUNCH_CHECK(ctx);
            uint32_t *in_id = pid, *out_id = pid2;
            XYZZ *in_val = pval, *out_val = pval2;
            while (nrec > 2048) {
                uint64_t nt = (nrec + COMBINE_LR - 1) / COMBINE_LR;
                msm_combine_level<<<(uint32_t)((nt + 127) / 128), 128, 0, st>>>(in_id, in_val, nrec, buckets, out_id, out_val, nt, giant_flag);
                B2_LAUNCH_CHECK(ctx);
                nrec = 2 * nt;
                uint32_t* ti = in_id; in_id = out_id; out_id = ti;
                XYZZ* tv = in_val; in_val = out_val; out_val = tv;
            }
            msm_combine_final<<<(uint32_t)((nrec + 127) / 128), 128, 0, st>>>(in_id, in_val, nrec, buckets, giant_flag);
            B2_LAUNCH_CHECK(ctx);
        }
    }
    {
        ProfScope ps_(ctx, PROF_MSM_REDUCE);
        msm_rowcol_sums<<<dim3(red_rows + red_cols, sets), RED_T, 0, st>>>(buckets, pl.B, kc, grpR);
        B2_LAUNCH_CHECK(ctx);
        msm_bit_sums<<<dim3(q_max + 1, 2, sets), RED_T, 0, st>>>(grpR, pl.B, kc, q_max, red_sums);
        B2_LAUNCH_CHECK(ctx);
        msm_finish<<<batch, 64, 0, st>>>(pl, kc, q_max, red_sums, out_dev);
        B2_LAUNCH_CHECK(ctx);
    }
    return B200ZK_OK;
}

int32_t msm_run(b200zk_ctx* ctx, const Affine* bases, const Fr* scalars, uint64_t n, Jacobian* out_dev, uint32_t pre_c,
                uint64_t pre_stride) {
    return msm_run_batch(ctx, bases, &scalars, 1, n, out_dev, pre_c, pre_stride);
}

int32_t g1_sum_run(b200zk_ctx* ctx, const Jacobian* pts, uint64_t count, Jacobian* out_dev) {
    g1_sum_kernel<<<1, 32, 0, ctx->stream>>>(pts, count, out_dev);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

int32_t g1_generator_mul_run(b200zk_ctx* ctx, const Fr* scalars, uint64_t n, Affine* out) {
    if (!n) return B200ZK_OK;
    g1_generator_mul_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(scalars, n, out);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

}  // namespace b200zk
