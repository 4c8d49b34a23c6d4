// Polynomial batch operations over BN254 Fr (HBM-bound streams, 32 B / element).
//
// Device replacements for the `parallelize(&mut [F], ..)` loops of halo2_proofs (Polynomial + - * scalar,
// pointwise products), arithmetic::eval_polynomial, arithmetic::kate_division and ff::BatchInvert
// (halo2_proofs/src/arithmetic.rs, src/poly.rs @ e5ddf67, pin /root/reference/Cargo.lock:1886-1888).
// Every thread moves whole 32 B elements as two 128-bit accesses; consecutive threads touch
// consecutive elements, grids are sized in multiples of the SM count.
#include "common.cuh"

namespace b200zk {

__device__ __forceinline__ Fr pl_ld(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fr r;
    r.l.v[0] = a.x; r.l.v[1] = a.y; r.l.v[2] = a.z; r.l.v[3] = a.w;
    r.l.v[4] = b.x; r.l.v[5] = b.y; r.l.v[6] = b.z; r.l.v[7] = b.w;
    return r;
}
__device__ __forceinline__ void pl_st(Fr* p, const Fr& r) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(r.l.v[0], r.l.v[1], r.l.v[2], r.l.v[3]);
    q[1] = make_uint4(r.l.v[4], r.l.v[5], r.l.v[6], r.l.v[7]);
}

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_SCALE = 3, OP_AXPY = 4 };

template <int OP>
__global__ void __launch_bounds__(256) poly_ew_kernel(Fr* r, const Fr* a, const Fr* b, Fr s, uint64_t n) {
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        Fr x = pl_ld(a + i), o;
        if (OP == OP_ADD) o = x + pl_ld(b + i);
        else if (OP == OP_SUB) o = x - pl_ld(b + i);
        else if (OP == OP_MUL) o = x * pl_ld(b + i);
        else if (OP == OP_SCALE) o = x * s;
        else o = x * s + pl_ld(b + i);
        pl_st(r + i, o);
    }
}

static uint32_t ew_blocks(b200zk_ctx* ctx, uint64_t n) {
    uint64_t want = (n + 255) / 256;
    uint64_t cap = (uint64_t)ctx->sm_count * 16;
    if (want > cap) want = cap;
    return (uint32_t)(want ? want : 1);
}

int32_t poly_ew(b200zk_ctx* ctx, int op, Fr* r, const Fr* a, const Fr* b, const Fr& s, uint64_t n) {
    if (n == 0) return B200ZK_OK;
    uint32_t blocks = ew_blocks(ctx, n);
    ProfScope ps_(ctx, PROF_POLY);
    switch (op) {
        case OP_ADD: poly_ew_kernel<OP_ADD><<<blocks, 256, 0, ctx->stream>>>(r, a, b, s, n); break;
        case OP_SUB: poly_ew_kernel<OP_SUB><<<blocks, 256, 0, ctx->stream>>>(r, a, b, s, n); break;
        case OP_MUL: poly_ew_kernel<OP_MUL><<<blocks, 256, 0, ctx->stream>>>(r, a, b, s, n); break;
        case OP_SCALE: poly_ew_kernel<OP_SCALE><<<blocks, 256, 0, ctx->stream>>>(r, a, b, s, n); break;
        default: poly_ew_kernel<OP_AXPY><<<blocks, 256, 0, ctx->stream>>>(r, a, b, s, n); break;
    }
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

// ---- eval_polynomial: thread t owns coefficients i = t (mod T); p_t = Horner in x^T; sum_t p_t x^t
__device__ __forceinline__ Fr block_sum(Fr v, Fr* sh) {  // sh: blockDim.x entries
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t s = blockDim.x >> 1; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + s];
        __syncthreads();
    }
    return sh[0];
}

__global__ void __launch_bounds__(256) eval_poly_kernel(const Fr* poly, uint64_t n, Fr x, Fr xT, uint32_t T, Fr* partial) {
    __shared__ Fr sh[256];
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    Fr acc = Fr::zero();
    if (t < T && t < n) {
        uint64_t cnt = (n - t + T - 1) / T;  // number of coefficients this thread owns
        for (uint64_t j = cnt; j-- > 0;) acc = acc * xT + pl_ld(poly + t + j * T);
        acc = acc * x.pow_u64(t);
    }
    Fr s = block_sum(acc, sh);
    if (threadIdx.x == 0) pl_st(partial + blockIdx.x, s);
}

__global__ void __launch_bounds__(256) sum_fr_kernel(const Fr* v, uint32_t cnt, Fr* out) {
    __shared__ Fr sh[256];
    Fr acc = Fr::zero();
    for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) acc = acc + pl_ld(v + i);
    Fr s = block_sum(acc, sh);
    if (threadIdx.x == 0) pl_st(out, s);
}

int32_t eval_poly(b200zk_ctx* ctx, const Fr* poly, uint64_t n, const Fr& x, Fr* out_dev) {
    uint32_t blocks = (uint32_t)ctx->sm_count * 2;
    if ((uint64_t)blocks * 256 > n) blocks = (uint32_t)((n + 255) / 256);
    if (blocks == 0) blocks = 1;
    uint32_t T = blocks * 256;
    Fr xT = x.pow_u64(T);
    B2_TRY(scratch_reserve(ctx, ctx->misc, sizeof(Fr) * (blocks + 1)));
    Fr* partial = (Fr*)ctx->misc.p;
    eval_poly_kernel<<<blocks, 256, 0, ctx->stream>>>(poly, n, x, xT, T, partial);
    B2_LAUNCH_CHECK(ctx);
    sum_fr_kernel<<<1, 256, 0, ctx->stream>>>(partial, blocks, out_dev);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

// ---- compute_inner_product(a, b) = sum_i a_i b_i
__global__ void __launch_bounds__(256) inner_product_kernel(const Fr* a, const Fr* b, uint64_t n, Fr* partial) {
    __shared__ Fr sh[256];
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    Fr acc = Fr::zero();
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) acc = acc + pl_ld(a + i) * pl_ld(b + i);
    Fr s = block_sum(acc, sh);
    if (threadIdx.x == 0) pl_st(partial + blockIdx.x, s);
}

int32_t inner_product(b200zk_ctx* ctx, const Fr* a, const Fr* b, uint64_t n, Fr* out_dev) {
    uint32_t blocks = ew_blocks(ctx, n);
    B2_TRY(scratch_reserve(ctx, ctx->misc, sizeof(Fr) * (blocks + 1)));
    Fr* partial = (Fr*)ctx->misc.p;
    inner_product_kernel<<<blocks, 256, 0, ctx->stream>>>(a, b, n, partial);
    B2_LAUNCH_CHECK(ctx);
    sum_fr_kernel<<<1, 256, 0, ctx->stream>>>(partial, blocks, out_dev);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

// ---- batch inversion: Montgomery's trick, hierarchical.  Thread t owns the strided slice {t, t+T, ...} (coalesced
// across the warp): pass 1 leaves the running products in `prefix` and the slice total in totals[t]; the totals are
// inverted by the same routine one level up (so the whole batch costs ONE Fermat inversion per 2048 leaf elements and
// 3 multiplications per element, instead of one inversion per 64 elements); pass 2 walks the slice backwards.  Zeros
// stay zero (they are skipped in the products), as ff::BatchInvert leaves them.
constexpr uint64_t BI_SLICE = 64, BI_LEAF = 2048;

__global__ void __launch_bounds__(256) batch_invert_up(const Fr* data, Fr* prefix, Fr* totals, uint64_t n, uint32_t T) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    Fr acc = Fr::one();
    for (uint64_t i = t; i < n; i += T) {
        pl_st(prefix + i, acc);
        Fr v = pl_ld(data + i);
        if (!v.is_zero()) acc = acc * v;
    }
    pl_st(totals + t, acc);
}
__global__ void __launch_bounds__(256) batch_invert_down(Fr* data, const Fr* prefix, const Fr* inv_totals, uint64_t n, uint32_t T) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T || t >= n) return;
    uint64_t cnt = (n - t + T - 1) / T;
    Fr acc = pl_ld(inv_totals + t);
    for (uint64_t j = cnt; j-- > 0;) {
        uint64_t i = t + j * T;
        Fr v = pl_ld(data + i);
        if (v.is_zero()) continue;
        pl_st(data + i, pl_ld(prefix + i) * acc);
        acc = acc * v;
    }
}
__global__ void __launch_bounds__(128) batch_invert_leaf(Fr* data, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pl_st(data + i, pl_ld(data + i).inv());  // 0 -> 0
}

// scratch: n + T + (next levels) elements
static int32_t batch_invert_level(b200zk_ctx* ctx, Fr* data, uint64_t n, Fr* scratch) {
    if (n <= BI_LEAF) {
        batch_invert_leaf<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(data, n);
        B2_LAUNCH_CHECK(ctx);
        return B200ZK_OK;
    }
    uint64_t T = (n + BI_SLICE - 1) / BI_SLICE;
    uint64_t cap = (uint64_t)ctx->sm_count * 2048;
    if (T > cap) T = cap;
    Fr *prefix = scratch, *totals = scratch + n;
    uint32_t blocks = (uint32_t)((T + 255) / 256);
    batch_invert_up<<<blocks, 256, 0, ctx->stream>>>(data, prefix, totals, n, (uint32_t)T);
    B2_LAUNCH_CHECK(ctx);
    B2_TRY(batch_invert_level(ctx, totals, T, totals + T));
    batch_invert_down<<<blocks, 256, 0, ctx->stream>>>(data, prefix, totals, n, (uint32_t)T);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

int32_t batch_invert(b200zk_ctx* ctx, Fr* data, uint64_t n) {
    if (n == 0) return B200ZK_OK;
    size_t elems = 0;
    for (uint64_t m = n; m > BI_LEAF;) {  // scratch of every level: prefix (m) + totals (T)
        uint64_t T = (m + BI_SLICE - 1) / BI_SLICE, cap = (uint64_t)ctx->sm_count * 2048;
        if (T > cap) T = cap;
        elems += m + T;
        m = T;
    }
    B2_TRY(scratch_reserve(ctx, ctx->misc, sizeof(Fr) * (elems + 1)));
    ProfScope ps_(ctx, PROF_POLY);
    return batch_invert_level(ctx, data, n, (Fr*)ctx->misc.p);
}

// ---- kate_division: q[j] = a[j+1] + b q[j+1], q[n-1] := 0  (three-phase chunked linear recurrence)
__global__ void __launch_bounds__(128) kate_phase1(const Fr* a, uint64_t nq, Fr b, uint64_t chunk, uint32_t nchunks, Fr* loc) {
    uint32_t cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= nchunks) return;
    uint64_t s = (uint64_t)cidx * chunk, e = s + chunk;
    if (e > nq) e = nq;
    Fr x = Fr::zero();
    for (uint64_t j = e; j-- > s;) x = pl_ld(a + j + 1) + b * x;
    pl_st(loc + cidx, x);
}
__global__ void kate_phase2(const Fr* loc, Fr b, uint64_t chunk, uint64_t nq, uint32_t nchunks, Fr* carry) {
    if (threadIdx.x || blockIdx.x) return;
    Fr bc = b.pow_u64(chunk);
    Fr c = Fr::zero();
    for (uint32_t k = nchunks; k-- > 0;) {
        pl_st(carry + k, c);  // carry-in of chunk k = q[end of chunk k]
        uint64_t s = (uint64_t)k * chunk, e = s + chunk;
        Fr f = (e > nq) ? b.pow_u64(nq - s) : bc;
        c = pl_ld(loc + k) + f * c;
    }
}
__global__ void __launch_bounds__(128) kate_phase3(const Fr* a, Fr* q, uint64_t nq, Fr b, uint64_t chunk, uint32_t nchunks,
                                                   const Fr* carry) {
    uint32_t cidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (cidx >= nchunks) return;
    uint64_t s = (uint64_t)cidx * chunk, e = s + chunk;
    if (e > nq) e = nq;
    Fr x = pl_ld(carry + cidx);
    for (uint64_t j = e; j-- > s;) {
        x = pl_ld(a + j + 1) + b * x;
        pl_st(q + j, x);
    }
}

int32_t kate_division(b200zk_ctx* ctx, Fr* q, const Fr* a, uint64_t n, const Fr& b) {
    if (n < 2) return B200ZK_OK;
    uint64_t nq = n - 1;
    uint64_t target = (uint64_t)ctx->sm_count * 64;  // chunks
    uint64_t chunk = (nq + target - 1) / target;
    if (chunk < 16) chunk = 16;
    uint32_t nchunks = (uint32_t)((nq + chunk - 1) / chunk);
    B2_TRY(scratch_reserve(ctx, ctx->misc, sizeof(Fr) * 2 * (size_t)nchunks));
    Fr* loc = (Fr*)ctx->misc.p;
    Fr* carry = loc + nchunks;
    uint32_t blocks = (nchunks + 127) / 128;
    kate_phase1<<<blocks, 128, 0, ctx->stream>>>(a, nq, b, chunk, nchunks, loc);
    B2_LAUNCH_CHECK(ctx);
    kate_phase2<<<1, 32, 0, ctx->stream>>>(loc, b, chunk, nq, nchunks, carry);
    B2_LAUNCH_CHECK(ctx);
    kate_phase3<<<blocks, 128, 0, ctx->stream>>>(a, q, nq, b, chunk, nchunks, carry);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

// ---- field-layer diagnostics
template <class F>
__global__ void field_op_kernel(int op, F* r, const F* a, const F* b, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i], o;
    switch (op) {
        case 0: o = x * y; break;
        case 1: o = x + y; break;
        case 2: o = x - y; break;
        case 3: o = x.inv(); break;
        case 4: o = x.from_mont(); break;
        default: o = x.sqr_scan(); break;
    }
    r[i] = o;
}

int32_t field_op(b200zk_ctx* ctx, int field, int op, void* r, const void* a, const void* b, uint64_t n) {
    if (n == 0) return B200ZK_OK;
    uint32_t blocks = (uint32_t)((n + 127) / 128);
    if (field == 0)
        field_op_kernel<Fr><<<blocks, 128, 0, ctx->stream>>>(op, (Fr*)r, (const Fr*)a, (const Fr*)b, n);
    else
        field_op_kernel<Fq><<<blocks, 128, 0, ctx->stream>>>(op, (Fq*)r, (const Fq*)a, (const Fq*)b, n);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

}  // namespace b200zk
