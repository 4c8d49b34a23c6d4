// Microbenchmarks that establish the INT32-multiply roofline denominators on this B200
// ("must be replaced by a measured IMAD microbenchmark").
//   imad_lo     : independent 32-bit IMAD chains              -> IMAD/s
//   imad_wide   : independent IMAD.WIDE.U32 (32x32+64) chains -> MAC32/s
//   fr_mul      : ILP independent Montgomery products/thread  -> MODMUL/s
//   fr_bfly     : t = b*w; (a+t, a-t)                         -> butterflies/s
// Prints one JSON line per measurement.  Build: see Makefile in this directory.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../scroll-prover_b200/csrc/ff.cuh"
using namespace b200zk;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_imad_lo(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(acc[i]) : "r"(a), "r"(b));
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_imad_wide(uint64_t* out, uint32_t a, uint32_t b, int iters) {
    uint64_t acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                uint32_t lo = (uint32_t)acc[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(lo ^ a), "r"(b));
            }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 pipe: independent DFMA chains (B200 keeps full-rate FP64, unlike B300) -- experiment for a 52-bit-limb
// floating-point modular multiplier running beside the INT32 pipe
template <int ILP>
__global__ void k_dfma(double* out, double a, double b, int iters) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// both pipes at once: one IMAD.WIDE-bound Montgomery product chain + independent DFMA chains in the same thread
template <int NDF>
__global__ void k_mix(Fr* data, double* out, double a, double b, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    Fr y = data[t], x = y;
    x.l.v[7] &= 0x0fffffff;
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x + i;
    for (int k = 0; k < iters; ++k) {
        x = x * y;
#pragma unroll
        for (int u = 0; u < NDF; ++u) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(acc[u & 7]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    data[t] = x;
    out[t] = s;
}

template <class F, int ILP>
__global__ void k_fmul(F* data, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    F x[ILP];
    F y = data[t];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { x[i] = y; x[i].l.v[0] ^= i; x[i].l.v[7] &= 0x0fffffff; }
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) x[i] = x[i] * y;
    }
    F s = x[0];
#pragma unroll
    for (int i = 1; i < ILP; ++i) s = s + x[i];
    data[t] = s;
}

template <class F, int ILP>
__global__ void k_bfly(F* data, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    F a[ILP], b[ILP];
    F w = data[t];
    w.l.v[7] &= 0x0fffffff;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = w; a[i].l.v[0] ^= i; b[i] = w; b[i].l.v[1] ^= i; }
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            F tt = b[i] * w;
            b[i] = a[i] - tt;
            a[i] = a[i] + tt;
        }
    }
    F s = a[0] + b[0];
#pragma unroll
    for (int i = 1; i < ILP; ++i) s = s + a[i] + b[i];
    data[t] = s;
}

template <class K>
static double time_ms(K launch, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    void* buf; CK(cudaMalloc(&buf, 64ull << 20));
    CK(cudaMemset(buf, 0x5a, 64ull << 20));
    const int iters = 2000;
    for (int tpb : {128, 256, 512, 1024}) {
        for (int bps : {1, 2, 4}) {
            if (tpb * bps > 2048) continue;
            int blocks = sms * bps;
            double ms = time_ms([&] { k_imad_lo<8><<<blocks, tpb>>>((uint32_t*)buf, 3, 5, iters); });
            double ops = (double)blocks * tpb * iters * 8 * 8;
            printf("{\"bench\": \"imad_lo\", \"tpb\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"Gops\": %.1f}\n", tpb, bps, ms, ops / ms / 1e6);
            ms = time_ms([&] { k_imad_wide<8><<<blocks, tpb>>>((uint64_t*)buf, 3, 5, iters); });
            printf("{\"bench\": \"imad_wide\", \"tpb\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"Gmac32\": %.1f}\n", tpb, bps, ms, ops / ms / 1e6);
        }
    }
    for (int bps : {2, 4, 8}) {
        int blocks = sms * bps, tpb = 256;
        double ms = time_ms([&] { k_dfma<8><<<blocks, tpb>>>((double*)buf, 1.0000001, 0.5, iters); });
        double ops = (double)blocks * tpb * iters * 8 * 8;
        printf("{\"bench\": \"dfma\", \"tpb\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"Gdfma\": %.1f}\n", tpb, bps, ms, ops / ms / 1e6);
    }
    {
        int blocks = sms * 4, tpb = 256, it = 500;
        double* dout = (double*)((char*)buf + (32ull << 20));
        double ms0 = time_ms([&] { k_mix<0><<<blocks, tpb>>>((Fr*)buf, dout, 1.0000001, 0.5, it); });
        double ms64 = time_ms([&] { k_mix<64><<<blocks, tpb>>>((Fr*)buf, dout, 1.0000001, 0.5, it); });
        double ms128 = time_ms([&] { k_mix<128><<<blocks, tpb>>>((Fr*)buf, dout, 1.0000001, 0.5, it); });
        double muls = (double)blocks * tpb * it;
        printf("{\"bench\": \"mix_frmul_plus_dfma\", \"ms_mul_only\": %.3f, \"ms_plus64dfma\": %.3f, \"ms_plus128dfma\": %.3f, \"Gmul_only\": %.2f, \"Gmul_with128\": %.2f, \"Gdfma_with128\": %.1f}\n",
               ms0, ms64, ms128, muls / ms0 / 1e6, muls / ms128 / 1e6, muls * 128 / ms128 / 1e6);
    }
    const int miters = 500;
#define RUN_F(NAME, KERN, ILP, TPB, BPS)                                                                       \
    {                                                                                                          \
        int blocks = sms * BPS;                                                                                \
        double ms = time_ms([&] { KERN<Fr, ILP><<<blocks, TPB>>>((Fr*)buf, miters); });                        \
        double ops = (double)blocks * TPB * miters * ILP;                                                      \
        printf("{\"bench\": \"%s\", \"ilp\": %d, \"tpb\": %d, \"blocks_per_sm\": %d, \"ms\": %.3f, \"Gops\": %.2f}\n", \
               NAME, ILP, TPB, BPS, ms, ops / ms / 1e6);                                                       \
    }
    RUN_F("fr_mul", k_fmul, 1, 256, 1) RUN_F("fr_mul", k_fmul, 1, 256, 2) RUN_F("fr_mul", k_fmul, 1, 256, 4)
    RUN_F("fr_mul", k_fmul, 1, 256, 6) RUN_F("fr_mul", k_fmul, 1, 512, 4)
    RUN_F("fr_mul", k_fmul, 2, 256, 1) RUN_F("fr_mul", k_fmul, 2, 256, 2) RUN_F("fr_mul", k_fmul, 2, 256, 4)
    RUN_F("fr_mul", k_fmul, 4, 128, 2) RUN_F("fr_mul", k_fmul, 4, 256, 1) RUN_F("fr_mul", k_fmul, 4, 256, 2) RUN_F("fr_mul", k_fmul, 4, 256, 3)
    RUN_F("fr_bfly", k_bfly, 1, 256, 4) RUN_F("fr_bfly", k_bfly, 2, 256, 2) RUN_F("fr_bfly", k_bfly, 2, 256, 4)
    RUN_F("fr_bfly", k_bfly, 4, 256, 1) RUN_F("fr_bfly", k_bfly, 4, 256, 2)
    // sustained: 2 s of fr_mul to see the clock the part settles at
    {
        int blocks = sms * 4;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int launches = 0;
        for (; launches < 400; ++launches) k_fmul<Fr, 2><<<blocks, 256>>>((Fr*)buf, miters);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)blocks * 256 * miters * 2 * launches;
        printf("{\"bench\": \"fr_mul_sustained\", \"ms\": %.1f, \"Gops\": %.2f}\n", ms, ops / ms / 1e6);
    }
    return 0;
}
