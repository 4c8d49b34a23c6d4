// FFT over BN254 G1 points (FftGroup = G1): device replacement for halo2_proofs::arithmetic::best_fft::<Fr, G1>
// and poly::kzg::commitment::g_to_lagrange, which Params::downsize uses to rebuild g_lagrange after truncating g
// (halo2_proofs/src/arithmetic.rs, src/poly/kzg/commitment.rs @ e5ddf67, pin /root/reference/Cargo.lock:1886-1888;
// reference call site /root/reference/integration/tests/integration.rs:17-18, README.md:22 "may affect performance").
//
// Radix-2 DIT on an XYZZ work array in HBM: bit-reversal on load, one kernel per stage, each thread one butterfly
// (t = w * b by double-and-add on the canonical twiddle bits; a' = a + t, b' = a - t), twiddles from the same
// universal per-stage table as the scalar NTT.  This is one-time SRS tooling, IMAD-bound, ~3300 MODMUL per butterfly (fixed signed 4-bit windows).
#include "common.cuh"
#include "ec.cuh"

namespace b200zk {

__device__ __forceinline__ XYZZ xyzz_neg(const XYZZ& p) {
    XYZZ r = p;
    r.y = p.y.neg();
    return r;
}

// s * P with fixed 4-bit SIGNED windows: s = sum_i d_i 16^i, d_i in [-8, 8]; table j*P, j = 1..8 (4 doublings + 3 additions),
// then 63 x 4 doublings + 64 additions -- no data-dependent branching on scalar bits, so the lanes of a warp (which all hold
// different twiddles) stay in lock step: ~3300 field multiplications instead of ~5500 for the bit-serial double-and-add
// whose `if (bit)` addition every lane ends up waiting for.
__device__ __noinline__ XYZZ xyzz_scalar_mul(const XYZZ& p, const Fr& s_mont) {
    Fr s = s_mont.from_mont();
    XYZZ T[8];
    T[0] = p;
    T[1] = xyzz_dbl(p);
    T[2] = T[1]; xyzz_add(T[2], p);
    T[3] = xyzz_dbl(T[1]);
    T[4] = T[3]; xyzz_add(T[4], p);
    T[5] = xyzz_dbl(T[2]);
    T[6] = T[5]; xyzz_add(T[6], p);
    T[7] = xyzz_dbl(T[3]);
    // digits from the least significant nibble up (carry = 1 when a nibble > 8 became negative); 254 bits -> 64 nibbles, the
    // top nibble is <= 3 so the last carry is absorbed
    int8_t dig[64];
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        uint32_t v = ((s.l.v[i >> 3] >> ((i & 7) * 4)) & 15u) + carry;
        carry = v > 8u;
        dig[i] = (int8_t)(carry ? (int)v - 16 : (int)v);
    }
    XYZZ acc = XYZZ::identity();
    for (int i = 63; i >= 0; --i) {
        if (i != 63) {
            acc = xyzz_dbl(acc);
            acc = xyzz_dbl(acc);
            acc = xyzz_dbl(acc);
            acc = xyzz_dbl(acc);
        }
        int d = dig[i];
        if (d != 0) {
            XYZZ t = T[(d < 0 ? -d : d) - 1];
            if (d < 0) t.y = t.y.neg();
            xyzz_add(acc, t);
        }
    }
    return acc;
}

template <bool FROM_JACOBIAN>
__global__ void __launch_bounds__(128) g1fft_load(const void* in, XYZZ* work, uint32_t log_n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ull << log_n)) return;
    uint64_t r = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
    XYZZ v = FROM_JACOBIAN ? xyzz_from_jacobian(((const Jacobian*)in)[i]) : xyzz_from_affine(((const Affine*)in)[i]);
    work[r] = v;
}

__global__ void __launch_bounds__(128) g1fft_stage(XYZZ* work, const Fr* __restrict__ tab, uint32_t log_n, uint32_t s) {
    uint64_t bb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (bb >= (1ull << (log_n - 1))) return;
    uint64_t half = 1ull << (s - 1);
    uint64_t K = bb & (half - 1), blk = bb >> (s - 1);
    uint64_t p0 = (blk << s) + K, p1 = p0 + half;
    XYZZ u = work[p0], v = work[p1];
    if (K != 0) v = xyzz_scalar_mul(v, tab[half + K]);
    XYZZ a = u, b = u;
    xyzz_add(a, v);
    xyzz_add(b, xyzz_neg(v));
    work[p0] = a;
    work[p1] = b;
}

template <bool TO_JACOBIAN>
__global__ void __launch_bounds__(128) g1fft_store(const XYZZ* work, void* out, uint64_t n, Fr scale, int do_scale) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ v = work[i];
    if (do_scale) v = xyzz_scalar_mul(v, scale);
    if (TO_JACOBIAN)
        ((Jacobian*)out)[i] = xyzz_to_jacobian_normalized(v);
    else
        ((Affine*)out)[i] = xyzz_to_affine(v);
}

// in/out: device pointers.  from_jac / to_jac select the 96 B Jacobian or 64 B affine encodings.
int32_t g1_fft_run(b200zk_ctx* ctx, const void* in, bool from_jac, void* out, bool to_jac, uint32_t log_n, const Fr& omega,
                   const Fr* scale) {
    if (log_n > 28) return fail(ctx, B200ZK_E_INVALID, "g1_fft: log_n %u > 28", log_n);
    uint64_t n = 1ull << log_n;
    const Fr* tab = nullptr;
    if (log_n >= 1) B2_TRY(ntt_get_table(ctx, omega, log_n, &tab));
    B2_TRY(scratch_reserve(ctx, ctx->msm_work, sizeof(XYZZ) * n));
    XYZZ* work = (XYZZ*)ctx->msm_work.p;
    uint32_t blocks = (uint32_t)((n + 127) / 128);
    if (from_jac)
        g1fft_load<true><<<blocks, 128, 0, ctx->stream>>>(in, work, log_n);
    else
        g1fft_load<false><<<blocks, 128, 0, ctx->stream>>>(in, work, log_n);
    B2_LAUNCH_CHECK(ctx);
    for (uint32_t s = 1; s <= log_n; ++s) {
        uint32_t b2 = (uint32_t)(((n >> 1) + 127) / 128);
        g1fft_stage<<<b2 ? b2 : 1, 128, 0, ctx->stream>>>(work, tab, log_n, s);
        B2_LAUNCH_CHECK(ctx);
    }
    Fr sc = scale ? *scale : Fr::one();
    if (to_jac)
        g1fft_store<true><<<blocks, 128, 0, ctx->stream>>>(work, out, n, sc, scale != nullptr);
    else
        g1fft_store<false><<<blocks, 128, 0, ctx->stream>>>(work, out, n, sc, scale != nullptr);
    B2_LAUNCH_CHECK(ctx);
    return B200ZK_OK;
}

}  // namespace b200zk
