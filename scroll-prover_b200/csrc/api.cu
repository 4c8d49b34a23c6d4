// extern "C" boundary of libb200zk (include/b200zk.h): argument checking, host<->device staging,
// error codes.  No exceptions cross this file; every failure path sets b200zk_last_error.
#include <new>

#include "common.cuh"
#include "ec.cuh"

namespace b200zk {
int32_t msm_run(b200zk_ctx* ctx, const Affine* bases, const Fr* scalars, uint64_t n, Jacobian* out_dev, uint32_t pre_c,
                uint64_t pre_stride);
uint32_t msm_pick_window_precomputed(uint64_t n);
int32_t srs_precompute_run(b200zk_ctx* ctx, Affine* tables, uint64_t n, uint32_t c, uint32_t W);
int32_t g1_sum_run(b200zk_ctx* ctx, const Jacobian* pts, uint64_t count, Jacobian* out_dev);
int32_t g1_generator_mul_run(b200zk_ctx* ctx, const Fr* scalars, uint64_t n, Affine* out);
int32_t poly_ew(b200zk_ctx* ctx, int op, Fr* r, const Fr* a, const Fr* b, const Fr& s, uint64_t n);
int32_t eval_poly(b200zk_ctx* ctx, const Fr* poly, uint64_t n, const Fr& x, Fr* out_dev);
int32_t batch_invert(b200zk_ctx* ctx, Fr* data, uint64_t n);
int32_t kate_division(b200zk_ctx* ctx, Fr* q, const Fr* a, uint64_t n, const Fr& b);
int32_t field_op(b200zk_ctx* ctx, int field, int op, void* r, const void* a, const void* b, uint64_t n);
}  // namespace b200zk

using namespace b200zk;

#define CHECK_CTX(ctx) \
    if (!(ctx)) return B200ZK_E_INVALID

static int32_t read_fr(b200zk_ctx* ctx, const void* p, Fr* out) {
    if (!p) return fail(ctx, B200ZK_E_INVALID, "null field element pointer");
    if (is_device_ptr(p)) {
        B2_CUDA(ctx, cudaMemcpyAsync(out, p, sizeof(Fr), cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    } else {
        memcpy(out, p, sizeof(Fr));
    }
    uint32_t m[8], d[8];
    Fr::modulus(m);
    if (!leaf::sub8(d, out->l.v, m)) return fail(ctx, B200ZK_E_INVALID, "field element is not reduced (>= modulus)");
    return B200ZK_OK;
}

// result delivery: dev -> (host | device) pointer
static int32_t deliver(b200zk_ctx* ctx, void* dst, const void* dev_src, size_t bytes) {
    if (is_device_ptr(dst)) {
        if (dst != dev_src) B2_CUDA(ctx, cudaMemcpyAsync(dst, dev_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        return B200ZK_OK;
    }
    return d2h(ctx, dst, dev_src, bytes);
}

extern "C" {

int32_t b200zk_ctx_create(const int* devices, int n_devices, b200zk_ctx** out) {
    if (!out) return B200ZK_E_INVALID;
    *out = nullptr;
    if (n_devices != 1 && !(n_devices == 0 && devices == nullptr)) return B200ZK_E_UNSUPPORTED;
    int dev = (devices && n_devices == 1) ? devices[0] : 0;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        return B200ZK_E_CUDA;  // no CPU fallback by design
    }
    if (dev < 0 || dev >= count) return B200ZK_E_INVALID;
    if (cudaSetDevice(dev) != cudaSuccess) return B200ZK_E_CUDA;
    b200zk_ctx* ctx = new (std::nothrow) b200zk_ctx();
    if (!ctx) return B200ZK_E_OOM;
    ctx->device = dev;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return B200ZK_E_CUDA;
    }
    ctx->own_stream = true;
    if (cudaMalloc(&ctx->msm_adds_dev, 8) == cudaSuccess) cudaMemset(ctx->msm_adds_dev, 0, 8);
    else ctx->msm_adds_dev = nullptr;
    *out = ctx;
    return B200ZK_OK;
}

int32_t b200zk_ctx_destroy(b200zk_ctx* ctx) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (Scratch* s : {&ctx->ntt_work, &ctx->stage_in, &ctx->stage_out, &ctx->msm_work, &ctx->misc})
        if (s->p) cudaFree(s->p);
    for (auto& t : ctx->tables) cudaFree(t.dev);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->msm_adds_dev) cudaFree(ctx->msm_adds_dev);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return B200ZK_OK;
}

const char* b200zk_last_error(const b200zk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int32_t b200zk_ctx_set_stream(b200zk_ctx* ctx, void* cuda_stream) {
    CHECK_CTX(ctx);
    Guard g(ctx);
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (cuda_stream == nullptr) {
        if (!ctx->own_stream) {
            B2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
            ctx->own_stream = true;
        }
        return B200ZK_OK;
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return B200ZK_OK;
}

int32_t b200zk_ctx_synchronize(b200zk_ctx* ctx) {
    CHECK_CTX(ctx);
    Guard g(ctx);
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200ZK_OK;
}

int32_t b200zk_ctx_launch_count(const b200zk_ctx* ctx, uint64_t* out) {
    if (!ctx || !out) return B200ZK_E_INVALID;
    *out = ctx->launches;
    return B200ZK_OK;
}

// ---- buffers ---------------------------------------------------------------------------------
int32_t b200zk_buf_alloc(b200zk_ctx* ctx, uint64_t bytes, void** out_dev) {
    CHECK_CTX(ctx);
    if (!out_dev) return fail(ctx, B200ZK_E_INVALID, "buf_alloc: null out");
    Guard g(ctx);
    *out_dev = nullptr;
    cudaError_t e = cudaMalloc(out_dev, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(ctx, B200ZK_E_OOM, "buf_alloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    }
    return B200ZK_OK;
}
int32_t b200zk_buf_free(b200zk_ctx* ctx, void* dev) {
    CHECK_CTX(ctx);
    Guard g(ctx);
    if (!dev) return B200ZK_OK;
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    B2_CUDA(ctx, cudaFree(dev));
    return B200ZK_OK;
}
int32_t b200zk_buf_upload(b200zk_ctx* ctx, void* dev, const void* host, uint64_t bytes) {
    CHECK_CTX(ctx);
    if (bytes && (!dev || !host)) return fail(ctx, B200ZK_E_INVALID, "buf_upload: null pointer");
    Guard g(ctx);
    B2_TRY(h2d(ctx, dev, host, bytes));
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // caller may reuse / free the host buffer
    return B200ZK_OK;
}
int32_t b200zk_buf_download(b200zk_ctx* ctx, void* host, const void* dev, uint64_t bytes) {
    CHECK_CTX(ctx);
    if (bytes && (!dev || !host)) return fail(ctx, B200ZK_E_INVALID, "buf_download: null pointer");
    Guard g(ctx);
    return d2h(ctx, host, dev, bytes);
}

// ---- SRS ---------------------------------------------------------------------------------------
int32_t b200zk_srs_register(b200zk_ctx* ctx, const void* g1_affine, uint64_t n, uint32_t tag, b200zk_srs** out) {
    CHECK_CTX(ctx);
    if (!out || (n && !g1_affine)) return fail(ctx, B200ZK_E_INVALID, "srs_register: null pointer");
    if (n >= (1ull << 31)) return fail(ctx, B200ZK_E_UNSUPPORTED, "srs_register: n >= 2^31");
    Guard g(ctx);
    b200zk_srs* s = new (std::nothrow) b200zk_srs();
    if (!s) return fail(ctx, B200ZK_E_OOM, "srs_register: host OOM");
    s->ctx = ctx;
    s->n = n;
    s->tag = tag;
    s->dev_bases = nullptr;
    s->pre_c = 0;
    s->pre_W = 1;
    size_t bytes = sizeof(Affine) * (n ? n : 1);
    if (ctx->srs_precompute && n >= (1ull << 16)) {
        // keep 2^(c*w) * P_i for every window w: all windows then share ONE bucket set (no per-window reduction, no
        // Horner doublings) and a wider window pays off.  Costs W x the base storage; skipped when memory is short.
        uint32_t c = msm_pick_window_precomputed(n), W = 254 / c + 1;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && (double)bytes * W < 0.35 * (double)free_b) {
            s->pre_c = c;
            s->pre_W = W;
            bytes *= W;
        }
    }
    cudaError_t e = cudaMalloc(&s->dev_bases, bytes);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        delete s;
        return fail(ctx, B200ZK_E_OOM, "srs_register: cudaMalloc(%zu) failed", bytes);
    }
    cudaMemcpyKind kind = is_device_ptr(g1_affine) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    e = n ? cudaMemcpyAsync(s->dev_bases, g1_affine, sizeof(Affine) * n, kind, ctx->stream) : cudaSuccess;
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        cudaFree(s->dev_bases);
        delete s;
        return fail(ctx, B200ZK_E_CUDA, "srs_register: upload failed: %s", cudaGetErrorString(e));
    }
    if (s->pre_c) {
        int32_t rc = srs_precompute_run(ctx, (Affine*)s->dev_bases, n, s->pre_c, s->pre_W);
        if (rc == B200ZK_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = B200ZK_E_CUDA;
        if (rc != B200ZK_OK) {
            cudaFree(s->dev_bases);
            delete s;
            return fail(ctx, rc, "srs_register: precomputation failed");
        }
    }
    *out = s;
    return B200ZK_OK;
}
int32_t b200zk_srs_release(b200zk_ctx* ctx, b200zk_srs* srs) {
    CHECK_CTX(ctx);
    if (!srs) return B200ZK_OK;
    Guard g(ctx);
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (srs->dev_bases) cudaFree(srs->dev_bases);
    delete srs;
    return B200ZK_OK;
}
int32_t b200zk_srs_len(const b200zk_srs* srs, uint64_t* out) {
    if (!srs || !out) return B200ZK_E_INVALID;
    *out = srs->n;
    return B200ZK_OK;
}

// ---- MSM ---------------------------------------------------------------------------------------
static int32_t msm_common(b200zk_ctx* ctx, const Affine* bases_dev, const void* scalars, uint64_t n, void* out96,
                          uint32_t pre_c = 0, uint64_t pre_stride = 0) {
    const void* sc_dev = nullptr;
    if (n) B2_TRY(stage_in(ctx, ctx->stage_in, scalars, sizeof(Fr) * n, &sc_dev));
    B2_TRY(scratch_reserve(ctx, ctx->stage_out, 256));
    Jacobian* res = (Jacobian*)ctx->stage_out.p;
    B2_TRY(msm_run(ctx, bases_dev, (const Fr*)sc_dev, n, res, pre_c, pre_stride));
    return deliver(ctx, out96, res, sizeof(Jacobian));
}

int32_t b200zk_msm_g1(b200zk_ctx* ctx, const b200zk_srs* srs, const void* scalars, uint64_t n, void* out_jacobian96) {
    CHECK_CTX(ctx);
    if (!srs || !out_jacobian96 || (n && !scalars)) return fail(ctx, B200ZK_E_INVALID, "msm_g1: null pointer");
    if (srs->ctx != ctx) return fail(ctx, B200ZK_E_INVALID, "msm_g1: SRS belongs to another context");
    if (n > srs->n)
        return fail(ctx, B200ZK_E_INVALID, "msm_g1: %llu scalars but only %llu bases (assert_eq!(coeffs.len(), bases.len()))",
                    (unsigned long long)n, (unsigned long long)srs->n);
    Guard g(ctx);
    return msm_common(ctx, (const Affine*)srs->dev_bases, scalars, n, out_jacobian96, srs->pre_c, srs->n);
}

int32_t b200zk_msm_g1_bases(b200zk_ctx* ctx, const void* g1_affine, const void* scalars, uint64_t n, void* out_jacobian96) {
    CHECK_CTX(ctx);
    if (!out_jacobian96 || (n && (!scalars || !g1_affine))) return fail(ctx, B200ZK_E_INVALID, "msm_g1_bases: null pointer");
    Guard g(ctx);
    const void* b_dev = nullptr;
    if (n) B2_TRY(stage_in(ctx, ctx->misc, g1_affine, sizeof(Affine) * n, &b_dev));
    // NB: ctx->misc is not used by msm_run
    return msm_common(ctx, (const Affine*)b_dev, scalars, n, out_jacobian96);
}

int32_t b200zk_g1_sum(b200zk_ctx* ctx, const void* jacobian_points, uint64_t count, void* out_jacobian96) {
    CHECK_CTX(ctx);
    if (!out_jacobian96 || (count && !jacobian_points)) return fail(ctx, B200ZK_E_INVALID, "g1_sum: null pointer");
    Guard g(ctx);
    const void* p_dev = nullptr;
    if (count) B2_TRY(stage_in(ctx, ctx->stage_in, jacobian_points, sizeof(Jacobian) * count, &p_dev));
    B2_TRY(scratch_reserve(ctx, ctx->stage_out, 256));
    Jacobian* res = (Jacobian*)ctx->stage_out.p;
    B2_TRY(g1_sum_run(ctx, (const Jacobian*)p_dev, count, res));
    return deliver(ctx, out_jacobian96, res, sizeof(Jacobian));
}

int32_t b200zk_g1_generator_mul_batch(b200zk_ctx* ctx, const void* scalars, uint64_t n, void* out_affine) {
    CHECK_CTX(ctx);
    if (n && (!scalars || !out_affine)) return fail(ctx, B200ZK_E_INVALID, "g1_generator_mul_batch: null pointer");
    Guard g(ctx);
    if (!n) return B200ZK_OK;
    const void* sc_dev = nullptr;
    B2_TRY(stage_in(ctx, ctx->stage_in, scalars, sizeof(Fr) * n, &sc_dev));
    bool out_dev = is_device_ptr(out_affine);
    Affine* res = (Affine*)out_affine;
    if (!out_dev) {
        B2_TRY(scratch_reserve(ctx, ctx->stage_out, sizeof(Affine) * n));
        res = (Affine*)ctx->stage_out.p;
    }
    B2_TRY(g1_generator_mul_run(ctx, (const Fr*)sc_dev, n, res));
    if (!out_dev) return d2h(ctx, out_affine, res, sizeof(Affine) * n);
    return B200ZK_OK;
}

// ---- NTT ---------------------------------------------------------------------------------------
int32_t b200zk_ntt_fr_ext(b200zk_ctx* ctx, const void* in, uint32_t log_in, void* out, uint32_t log_n, const void* omega32,
                          int inverse_scale, int coset_mode) {
    CHECK_CTX(ctx);
    if (!in || !out || !omega32) return fail(ctx, B200ZK_E_INVALID, "ntt: null pointer");
    if (log_n > 28 || log_in > log_n) return fail(ctx, B200ZK_E_INVALID, "ntt: bad sizes log_in=%u log_n=%u", log_in, log_n);
    Guard g(ctx);
    Fr omega;
    B2_TRY(read_fr(ctx, omega32, &omega));
    size_t in_bytes = sizeof(Fr) << log_in, out_bytes = sizeof(Fr) << log_n;
    bool out_is_dev = is_device_ptr(out);
    Fr* out_dev = (Fr*)out;
    if (!out_is_dev) {
        B2_TRY(scratch_reserve(ctx, ctx->stage_out, out_bytes));
        out_dev = (Fr*)ctx->stage_out.p;
    }
    const void* in_dev = nullptr;
    if (!is_device_ptr(in) && !out_is_dev && log_in == log_n) {
        // host in / host out of equal size: upload straight into the output staging buffer and run in place
        B2_TRY(h2d(ctx, out_dev, in, in_bytes));
        in_dev = out_dev;
    } else {
        B2_TRY(stage_in(ctx, ctx->stage_in, in, in_bytes, &in_dev));
    }
    B2_TRY(ntt_run(ctx, (const Fr*)in_dev, log_in, out_dev, log_n, omega, inverse_scale, coset_mode));
    if (!out_is_dev) return d2h(ctx, out, out_dev, out_bytes);
    return B200ZK_OK;
}

int32_t b200zk_ntt_fr(b200zk_ctx* ctx, void* data, uint32_t log_n, const void* omega32, int inverse_scale, int coset_mode) {
    return b200zk_ntt_fr_ext(ctx, data, log_n, data, log_n, omega32, inverse_scale, coset_mode);
}

// ---- poly ops ------------------------------------------------------------------------------------
static int32_t ew_common(b200zk_ctx* ctx, int op, void* r, const void* a, const void* b, const void* s32, uint64_t n) {
    CHECK_CTX(ctx);
    bool need_b = (op == 0 || op == 1 || op == 2 || op == 4), need_s = (op == 3 || op == 4);
    if (n && (!r || !a || (need_b && !b) || (need_s && !s32))) return fail(ctx, B200ZK_E_INVALID, "poly op: null pointer");
    Guard g(ctx);
    if (!n) return B200ZK_OK;
    Fr s = Fr::zero();
    if (need_s) B2_TRY(read_fr(ctx, s32, &s));
    size_t bytes = sizeof(Fr) * n;
    const void *a_dev = nullptr, *b_dev = nullptr;
    B2_TRY(stage_in(ctx, ctx->stage_in, a, bytes, &a_dev));
    if (need_b) B2_TRY(stage_in(ctx, ctx->ntt_work, b, bytes, &b_dev));
    bool r_is_dev = is_device_ptr(r);
    Fr* r_dev = (Fr*)r;
    if (!r_is_dev) {
        B2_TRY(scratch_reserve(ctx, ctx->stage_out, bytes));
        r_dev = (Fr*)ctx->stage_out.p;
    }
    B2_TRY(poly_ew(ctx, op, r_dev, (const Fr*)a_dev, (const Fr*)b_dev, s, n));
    if (!r_is_dev) return d2h(ctx, r, r_dev, bytes);
    return B200ZK_OK;
}
int32_t b200zk_poly_add(b200zk_ctx* ctx, void* r, const void* a, const void* b, uint64_t n) { return ew_common(ctx, 0, r, a, b, nullptr, n); }
int32_t b200zk_poly_sub(b200zk_ctx* ctx, void* r, const void* a, const void* b, uint64_t n) { return ew_common(ctx, 1, r, a, b, nullptr, n); }
int32_t b200zk_poly_mul(b200zk_ctx* ctx, void* r, const void* a, const void* b, uint64_t n) { return ew_common(ctx, 2, r, a, b, nullptr, n); }
int32_t b200zk_poly_scale(b200zk_ctx* ctx, void* r, const void* a, const void* s32, uint64_t n) { return ew_common(ctx, 3, r, a, nullptr, s32, n); }
int32_t b200zk_poly_axpy(b200zk_ctx* ctx, void* r, const void* a, const void* s32, const void* b, uint64_t n) {
    return ew_common(ctx, 4, r, a, b, s32, n);
}

int32_t b200zk_eval_poly(b200zk_ctx* ctx, const void* poly, uint64_t n, const void* point32, void* out32) {
    CHECK_CTX(ctx);
    if (!out32 || !point32 || (n && !poly)) return fail(ctx, B200ZK_E_INVALID, "eval_poly: null pointer");
    Guard g(ctx);
    Fr x;
    B2_TRY(read_fr(ctx, point32, &x));
    B2_TRY(scratch_reserve(ctx, ctx->stage_out, 256));
    Fr* res = (Fr*)ctx->stage_out.p;
    if (!n) {
        B2_CUDA(ctx, cudaMemsetAsync(res, 0, sizeof(Fr), ctx->stream));
    } else {
        const void* p_dev = nullptr;
        B2_TRY(stage_in(ctx, ctx->stage_in, poly, sizeof(Fr) * n, &p_dev));
        B2_TRY(eval_poly(ctx, (const Fr*)p_dev, n, x, res));
    }
    return deliver(ctx, out32, res, sizeof(Fr));
}

int32_t b200zk_batch_invert(b200zk_ctx* ctx, void* data, uint64_t n) {
    CHECK_CTX(ctx);
    if (n && !data) return fail(ctx, B200ZK_E_INVALID, "batch_invert: null pointer");
    Guard g(ctx);
    if (!n) return B200ZK_OK;
    if (is_device_ptr(data)) return batch_invert(ctx, (Fr*)data, n);
    size_t bytes = sizeof(Fr) * n;
    B2_TRY(scratch_reserve(ctx, ctx->stage_out, bytes));
    B2_TRY(h2d(ctx, ctx->stage_out.p, data, bytes));
    B2_TRY(batch_invert(ctx, (Fr*)ctx->stage_out.p, n));
    return d2h(ctx, data, ctx->stage_out.p, bytes);
}

int32_t b200zk_kate_division(b200zk_ctx* ctx, void* q, const void* a, uint64_t n, const void* b32) {
    CHECK_CTX(ctx);
    if (n < 1 || !a || !b32 || (n > 1 && !q)) return fail(ctx, B200ZK_E_INVALID, "kate_division: bad arguments");
    Guard g(ctx);
    if (n == 1) return B200ZK_OK;
    Fr b;
    B2_TRY(read_fr(ctx, b32, &b));
    const void* a_dev = nullptr;
    B2_TRY(stage_in(ctx, ctx->stage_in, a, sizeof(Fr) * n, &a_dev));
    bool q_is_dev = is_device_ptr(q);
    Fr* q_dev = (Fr*)q;
    size_t qbytes = sizeof(Fr) * (n - 1);
    if (!q_is_dev) {
        B2_TRY(scratch_reserve(ctx, ctx->stage_out, qbytes));
        q_dev = (Fr*)ctx->stage_out.p;
    }
    B2_TRY(kate_division(ctx, q_dev, (const Fr*)a_dev, n, b));
    if (!q_is_dev) return d2h(ctx, q, q_dev, qbytes);
    return B200ZK_OK;
}

// ---- diagnostics -----------------------------------------------------------------------------------
int32_t b200zk_debug_field_op(b200zk_ctx* ctx, int field, int op, void* r, const void* a, const void* b, uint64_t n) {
    CHECK_CTX(ctx);
    if (n && (!r || !a || !b)) return fail(ctx, B200ZK_E_INVALID, "debug_field_op: null pointer");
    Guard g(ctx);
    if (!n) return B200ZK_OK;
    size_t bytes = 32 * n;
    const void *a_dev = nullptr, *b_dev = nullptr;
    B2_TRY(stage_in(ctx, ctx->stage_in, a, bytes, &a_dev));
    B2_TRY(stage_in(ctx, ctx->ntt_work, b, bytes, &b_dev));
    B2_TRY(scratch_reserve(ctx, ctx->stage_out, bytes));
    B2_TRY(field_op(ctx, field, op, ctx->stage_out.p, a_dev, b_dev, n));
    return deliver(ctx, r, ctx->stage_out.p, bytes);
}

int32_t b200zk_profile_enable(b200zk_ctx* ctx, int on) {
    CHECK_CTX(ctx);
    Guard g(ctx);
    ctx->profiling = on != 0;
    return B200ZK_OK;
}
int32_t b200zk_profile_reset(b200zk_ctx* ctx) {
    CHECK_CTX(ctx);
    Guard g(ctx);
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    prof_resolve(ctx);
    for (int i = 0; i < PROF_NKEYS; ++i) {
        ctx->prof_ms[i] = 0;
        ctx->prof_cnt[i] = 0;
    }
    return B200ZK_OK;
}
int32_t b200zk_profile_read(b200zk_ctx* ctx, const char* name, double* total_ms, uint64_t* count) {
    CHECK_CTX(ctx);
    if (!name) return fail(ctx, B200ZK_E_INVALID, "profile_read: null name");
    Guard g(ctx);
    B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    prof_resolve(ctx);
    for (int i = 0; i < PROF_NKEYS; ++i)
        if (strcmp(name, PROF_NAMES[i]) == 0) {
            if (total_ms) *total_ms = ctx->prof_ms[i];
            if (count) *count = ctx->prof_cnt[i];
            return B200ZK_OK;
        }
    return fail(ctx, B200ZK_E_INVALID, "profile_read: unknown kernel class '%s'", name);
}

int32_t b200zk_srs_set_precompute(b200zk_ctx* ctx, int mode) {
    CHECK_CTX(ctx);
    if (mode != 0 && mode != 1) return fail(ctx, B200ZK_E_INVALID, "srs_set_precompute: mode must be 0 or 1");
    ctx->srs_precompute = mode;
    return B200ZK_OK;
}
int32_t b200zk_msm_set_window(b200zk_ctx* ctx, uint32_t c) {
    CHECK_CTX(ctx);
    if (c != 0 && (c < 2 || c > 24)) return fail(ctx, B200ZK_E_INVALID, "msm window %u out of range", c);
    ctx->msm_window = c;
    return B200ZK_OK;
}
int32_t b200zk_msm_total_adds(b200zk_ctx* ctx, uint64_t* actual_adds, int reset) {
    CHECK_CTX(ctx);
    Guard g(ctx);
    unsigned long long v = 0;
    if (ctx->msm_adds_dev) {
        B2_CUDA(ctx, cudaMemcpyAsync(&v, ctx->msm_adds_dev, 8, cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (reset) B2_CUDA(ctx, cudaMemsetAsync(ctx->msm_adds_dev, 0, 8, ctx->stream));
    }
    if (actual_adds) *actual_adds = v;
    return B200ZK_OK;
}
int32_t b200zk_msm_last_stats(const b200zk_ctx* ctx, uint32_t* window_bits, uint32_t* n_windows, uint64_t* n_bucket_adds) {
    if (!ctx) return B200ZK_E_INVALID;
    if (window_bits) *window_bits = ctx->last_c;
    if (n_windows) *n_windows = ctx->last_windows;
    if (n_bucket_adds) *n_bucket_adds = ctx->last_adds;
    return B200ZK_OK;
}

}  // extern "C"
