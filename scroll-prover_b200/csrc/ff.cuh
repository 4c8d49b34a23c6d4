// BN254 prime fields on sm_100a: 8 x u32 limbs, Montgomery form R = 2^256.
//
// Device replacement for halo2curves::bn256::{Fr,Fq} (halo2curves 0.1.0 @ 112f5b9, pin
// /root/reference/Cargo.lock:1911-1913; src/bn256/{fr.rs,fq.rs}, src/derive/field.rs).  In-memory
// layout is identical to the Rust types (4 x u64 LE Montgomery limbs == 8 x u32 LE), so host
// buffers are memcpy-compatible and all results are the fully reduced representative in [0,p).
//
// Multiplication is word-serial CIOS on the INT32 multiply pipe: per multiplier word one
// product row (a*b_i) and one reduction row (m*p), each split into an even-limb and an
// odd-limb carry chain (mad.lo.cc / madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32
// with predicate carries).  Two accumulators, one aligned at limb 0 and one at limb 1,
// swap roles after each row so no limb shuffling is needed.  128 32x32->64 MACs per mul.
//
// Every carry chain lives inside ONE asm statement (the PTX CC flag never crosses a
// statement).  Under host compilation (no __CUDA_ARCH__) the same leaf primitives are
// emulated with 64-bit arithmetic, so the limb choreography is unit-testable on a CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FF_HD __host__ __device__ __forceinline__
#define FF_D __device__ __forceinline__
#else
#define FF_HD inline
#define FF_D inline
#endif

namespace b200zk {

struct alignas(16) limbs8 {
    uint32_t v[8];
};

// ---------------------------------------------------------------------------------------------
// field configurations
// ---------------------------------------------------------------------------------------------
struct FrCfg {  // scalar field r
    FF_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    FF_HD static constexpr uint32_t one(int i) {  // R mod r
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    FF_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xefffffffu;  // -r^-1 mod 2^32
};

struct FqCfg {  // base field q
    FF_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    FF_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    FF_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
    static constexpr uint32_t INV = 0xe4866389u;
};

// ---------------------------------------------------------------------------------------------
// leaf carry-chain primitives (device: PTX; host: emulation)
// ---------------------------------------------------------------------------------------------
namespace leaf {

// acc[j], acc[j+1] = lo, hi (a[j] * b), j = 0,2,4,6      (no carries: disjoint 64-bit products)
FF_HD void mul_even(uint32_t* acc, const uint32_t* a, uint32_t b) {
#ifdef __CUDA_ARCH__
#pragma unroll
    for (int j = 0; j < 8; j += 2)
        asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(b));
#else
    for (int j = 0; j < 8; j += 2) {
        uint64_t p = (uint64_t)a[j] * b;
        acc[j] = (uint32_t)p;
        acc[j + 1] = (uint32_t)(p >> 32);
    }
#endif
}

// acc[0..8) += sum_{j even} a[j]*b << (32 j); the carry out of limb 7 is added to *top.
FF_HD void cmad_even_top(uint32_t* acc, const uint32_t* a, uint32_t b, uint32_t& top) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
#else
    uint64_t c = 0;
    for (int j = 0; j < 8; j += 2) {
        uint64_t p = (uint64_t)a[j] * b;
        uint64_t s = (uint64_t)acc[j] + (uint32_t)p + c;
        acc[j] = (uint32_t)s;
        s = (uint64_t)acc[j + 1] + (uint32_t)(p >> 32) + (s >> 32);
        acc[j + 1] = (uint32_t)s;
        c = s >> 32;
    }
    top += (uint32_t)c;
#endif
}

// same, carry out of limb 7 dropped (caller guarantees it is zero)
FF_HD void cmad_even(uint32_t* acc, const uint32_t* a, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
#else
    uint32_t dummy = 0;
    cmad_even_top(acc, a, b, dummy);
#endif
}

// x0 += y[1] (carry c);  y[j],y[j+1] = a[j]*b + (y[j+2],y[j+3]) + c  for j = 0,2,4 ;
// y[6],y[7] = a[6]*b + c.     ("shift the odd accumulator down two limbs while accumulating")
FF_HD void shift_mad_even(uint32_t& x0, uint32_t* y, const uint32_t* a, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("add.cc.u32 %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
        "madc.hi.u32 %7, %12, %13, 0;"
        : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7]), "+r"(x0)
        : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(b));
#else
    uint64_t s = (uint64_t)x0 + y[1];
    x0 = (uint32_t)s;
    uint64_t c = s >> 32;
    for (int j = 0; j < 8; j += 2) {
        uint64_t p = (uint64_t)a[j] * b;
        uint32_t add_lo = (j < 6) ? y[j + 2] : 0u, add_hi = (j < 6) ? y[j + 3] : 0u;
        s = (uint64_t)add_lo + (uint32_t)p + c;
        y[j] = (uint32_t)s;
        s = (uint64_t)add_hi + (uint32_t)(p >> 32) + (s >> 32);
        y[j + 1] = (uint32_t)s;
        c = s >> 32;
    }
#endif
}

// (c2:c1:c0) += x * y      (96-bit column accumulator of the squaring's product scanning)
FF_HD void mac3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t x, uint32_t y) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32 %2, %2, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2)
        : "r"(x), "r"(y));
#else
    uint64_t p = (uint64_t)x * y;
    uint64_t s = (uint64_t)c0 + (uint32_t)p;
    c0 = (uint32_t)s;
    s = (uint64_t)c1 + (uint32_t)(p >> 32) + (s >> 32);
    c1 = (uint32_t)s;
    c2 += (uint32_t)(s >> 32);
#endif
}

// t[0..16) = d[0..16) + sum_i a[i]^2 << (64 i)    (one carry chain over all 16 limbs; no carry out: a^2 < 2^512)
FF_HD void add_diag(uint32_t* t, const uint32_t* d, const uint32_t* a) {
#ifdef __CUDA_ARCH__
    asm("mad.lo.cc.u32 %0, %16, %16, %24;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %25;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %26;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %27;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %28;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %29;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %30;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %31;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %32;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %33;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %34;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %35;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %36;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %37;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %38;\n\t"
        "madc.hi.u32 %15, %23, %23, %39;"
        : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]),
          "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(d[0]), "r"(d[1]), "r"(d[2]),
          "r"(d[3]), "r"(d[4]), "r"(d[5]), "r"(d[6]), "r"(d[7]), "r"(d[8]), "r"(d[9]), "r"(d[10]), "r"(d[11]), "r"(d[12]), "r"(d[13]),
          "r"(d[14]), "r"(d[15]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t p = (uint64_t)a[i] * a[i];
        uint64_t s = (uint64_t)d[2 * i] + (uint32_t)p + c;
        t[2 * i] = (uint32_t)s;
        s = (uint64_t)d[2 * i + 1] + (uint32_t)(p >> 32) + (s >> 32);
        t[2 * i + 1] = (uint32_t)s;
        c = s >> 32;
    }
#endif
}

// r = a + b (8 limbs), returns carry out
FF_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t s = (uint64_t)a[i] + b[i] + c;
        r[i] = (uint32_t)s;
        c = s >> 32;
    }
    return (uint32_t)c;
#endif
}

// r = a - b (8 limbs), returns borrow (1 if a < b)
FF_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#ifdef __CUDA_ARCH__
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;  // subc of 0-0-borrow = 0xffffffff when borrow
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; ++i) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    return (uint32_t)br;
#endif
}

}  // namespace leaf

// ---------------------------------------------------------------------------------------------
// field element
// ---------------------------------------------------------------------------------------------
template <class Cfg>
struct Fp {
    limbs8 l;

    FF_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l.v[i] = 0;
        return r;
    }
    FF_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l.v[i] = Cfg::one(i);
        return r;
    }
    FF_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l.v[i] = Cfg::r2(i);
        return r;
    }
    FF_HD static void modulus(uint32_t* m) {
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = Cfg::mod(i);
    }

    FF_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= l.v[i];
        return o == 0;
    }
    FF_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= l.v[i] ^ b.l.v[i];
        return o == 0;
    }
    FF_HD bool operator!=(const Fp& b) const { return !(*this == b); }

    // t in [0, 2p) -> [0, p)
    FF_HD static void final_sub(uint32_t* t) {
        uint32_t m[8], d[8];
        modulus(m);
        uint32_t borrow = leaf::sub8(d, t, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) t[i] = borrow ? t[i] : d[i];
    }

    FF_HD friend Fp operator+(const Fp& a, const Fp& b) {
        Fp r;
        leaf::add8(r.l.v, a.l.v, b.l.v);  // < 2p < 2^255: no carry out
        final_sub(r.l.v);
        return r;
    }
    FF_HD friend Fp operator-(const Fp& a, const Fp& b) {
        Fp r;
        uint32_t m[8], t[8];
        modulus(m);
        uint32_t borrow = leaf::sub8(r.l.v, a.l.v, b.l.v);
        leaf::add8(t, r.l.v, m);
#pragma unroll
        for (int i = 0; i < 8; ++i) r.l.v[i] = borrow ? t[i] : r.l.v[i];
        return r;
    }
    FF_HD Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    FF_HD Fp dbl() const { return *this + *this; }

    // Montgomery product a*b*R^-1 mod p, fully reduced.
    FF_HD friend Fp operator*(const Fp& a, const Fp& b) {
        uint32_t m[8];
        modulus(m);
        uint32_t even[8], odd[8];
        const uint32_t* av = a.l.v;
        // row 0
        leaf::mul_even(odd, av + 1, b.l.v[0]);
        leaf::mul_even(even, av, b.l.v[0]);
        {
            uint32_t mi = even[0] * Cfg::INV;
            leaf::cmad_even(odd, m + 1, mi);
            leaf::cmad_even_top(even, m, mi, odd[7]);
        }
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            // X = accumulator aligned at limb 0 for this row, Y = the other one
            uint32_t* X = (i & 1) ? odd : even;
            uint32_t* Y = (i & 1) ? even : odd;
            uint32_t bi = b.l.v[i];
            leaf::shift_mad_even(X[0], Y, av + 1, bi);
            leaf::cmad_even_top(X, av, bi, Y[7]);
            uint32_t mi = X[0] * Cfg::INV;
            leaf::cmad_even(Y, m + 1, mi);
            leaf::cmad_even_top(X, m, mi, Y[7]);
        }
        // after row 7: X = odd, Y = even;  result = even + (odd >> 32)
        Fp r;
        uint32_t sh[8];
#pragma unroll
        for (int i = 0; i < 7; ++i) sh[i] = odd[i + 1];
        sh[7] = 0;
        leaf::add8(r.l.v, even, sh);
        final_sub(r.l.v);
        return r;
    }
    // sqr(): the general product.  Measured on B200 (profiles/, round 1): the dedicated square below issues 100
    // instead of 128 IMAD.WIDE but its product-scanning column chain is serial, so msm_accumulate got 2 % slower and the
    // latency-bound bucket-reduction kernels do not want a longer dependency chain either.
    FF_HD Fp sqr() const { return (*this) * (*this); }

    // Dedicated Montgomery square (kept, tested, for throughput-bound tooling such as the fixed-base generator
    // multiplication): 36 product MACs (28 off-diagonal by product scanning, doubled, + 8 diagonal) followed by the 64
    // reduction MACs.  The carries that leave a reduction row land at limb >= 8 and never influence a later
    // Montgomery quotient, so they are collected and added once.
    FF_HD Fp sqr_scan() const {
        const uint32_t* a = l.v;
        uint32_t off[16], d[16], t[17];
        uint32_t c0 = 0, c1 = 0, c2 = 0;
        off[0] = 0;
#pragma unroll
        for (int k = 1; k <= 13; ++k) {
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int j = k - i;
                if (j > i && j <= 7) leaf::mac3(c0, c1, c2, a[i], a[j]);
            }
            off[k] = c0;
            c0 = c1;
            c1 = c2;
            c2 = 0;
        }
        off[14] = c0;
        off[15] = c1;
        d[0] = 0;
#pragma unroll
        for (int k = 1; k < 16; ++k) d[k] = (off[k] << 1) | (off[k - 1] >> 31);
        leaf::add_diag(t, d, a);
        t[16] = 0;
        uint32_t m[8], e[9];
        modulus(m);
#pragma unroll
        for (int i = 0; i < 9; ++i) e[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t mi = t[i] * Cfg::INV;
            leaf::cmad_even_top(t + i, m, mi, e[i]);          // limbs i..i+7, carry -> limb i+8
            leaf::cmad_even_top(t + i + 1, m + 1, mi, e[i + 1]);  // limbs i+1..i+8, carry -> limb i+9
        }
        Fp r;
        leaf::add8(r.l.v, t + 8, e);  // e[k] = carries into limb 8+k; e[8] (limb 16) is always 0: the result is < 2p
        final_sub(r.l.v);
        return r;
    }

    // canonical (non-Montgomery) limbs: to_repr()
    FF_HD Fp from_mont() const {
        Fp o = zero();
        o.l.v[0] = 1;
        return (*this) * o;
    }
    FF_HD Fp to_mont() const { return (*this) * r2(); }

    FF_HD Fp pow_u64(uint64_t e) const {
        Fp acc = one(), base = *this;
        while (e) {
            if (e & 1) acc = acc * base;
            base = base.sqr();
            e >>= 1;
        }
        return acc;
    }
    // a^(p-2); 0 -> 0
    FF_HD Fp inv() const {
        uint32_t e[8];
        modulus(e);
        e[0] -= 2;  // low limb of both moduli is >= 2
        Fp acc = one();
        for (int i = 7; i >= 0; --i)
            for (int b = 31; b >= 0; --b) {
                acc = acc.sqr();
                if ((e[i] >> b) & 1) acc = acc * (*this);
            }
        return acc;
    }
};

using Fr = Fp<FrCfg>;
using Fq = Fp<FqCfg>;

}  // namespace b200zk
