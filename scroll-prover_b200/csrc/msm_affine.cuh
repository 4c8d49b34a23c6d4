// Batched-affine bucket accumulation for the MSM (EXPERIMENTAL, off by default: B200ZK_MSM_AFFINE=1).
//
// msm_accumulate adds affine bases into XYZZ accumulators: 10 field multiplications per bucket addition, and it already
// runs at ~0.9 of the INT32-multiply roofline, so the only way down is fewer multiplications.  An affine + affine
// addition costs 3 (lambda, lambda^2, y3) plus one inversion; Montgomery's trick shares the inversion over many
// independent additions at 3 more multiplications each, i.e. ~6 per addition.  Independence comes from a pairwise tree
// over the SORTED entries of every bucket: level l+1 holds ceil(m/2) points per bucket (adjacent pairs summed, an odd
// leftover copied), so a bucket of m entries is done after ceil(log2 m) levels and the total number of additions is
// unchanged (entries - buckets).
//
// Each level is two streaming kernels around one small batch inversion:
//   A  thread t owns L consecutive OUTPUT slots: classifies each pair (add / double / cancel / copy), multiplies the
//      denominators into a running product, stores the product BEFORE each slot (prefix) and its total;
//      the totals (outputs / L of them) are inverted by batch_invert (poly.cu);
//   B  the same thread walks its slots backwards, peels one inverse per slot off the inverted total, and writes the
//      affine sums.
// Exceptional cases are decided from the inputs alone, identically in A and B: identity inputs (0,0) -> copy, equal
// points -> tangent (denominator 2y), opposite points -> identity, single leftover -> copy.
//
// This header holds the per-thread bodies as FF_HD functions of (thread index, arrays): the CUDA kernels in
// msm_affine.cu are one-line wrappers, and tests/test_msm_affine_host.py runs the very same bodies thread by thread on
// the CPU against straightforward bucket sums (the threads of a level never communicate, so that emulation is exact).
#pragma once
#include "ec.cuh"

namespace b200zk {

struct BaLevel {
    const Affine* bases;      // level 0: SRS (or its precomputed tables), addressed through `entries`
    const uint32_t* entries;  // level 0: sorted (table index | sign << 31); nullptr on inner levels
    const Affine* points;     // inner levels: the previous level's outputs
    const uint32_t* off_in;   // NB + 1 bucket offsets into entries / points
    const uint32_t* off_out;  // NB + 1 bucket offsets of this level's outputs (counts = ceil(m_in / 2))
    uint64_t NB;
};

FF_HD Affine ba_load(const Affine* p) {
#ifdef __CUDA_ARCH__
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1], c = q[2], d = q[3];
    Affine r;
    r.x.l.v[0] = a.x; r.x.l.v[1] = a.y; r.x.l.v[2] = a.z; r.x.l.v[3] = a.w;
    r.x.l.v[4] = b.x; r.x.l.v[5] = b.y; r.x.l.v[6] = b.z; r.x.l.v[7] = b.w;
    r.y.l.v[0] = c.x; r.y.l.v[1] = c.y; r.y.l.v[2] = c.z; r.y.l.v[3] = c.w;
    r.y.l.v[4] = d.x; r.y.l.v[5] = d.y; r.y.l.v[6] = d.z; r.y.l.v[7] = d.w;
    return r;
#else
    return *p;
#endif
}
FF_HD void ba_store(Affine* p, const Affine& v) {
#ifdef __CUDA_ARCH__
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.x.l.v[0], v.x.l.v[1], v.x.l.v[2], v.x.l.v[3]);
    q[1] = make_uint4(v.x.l.v[4], v.x.l.v[5], v.x.l.v[6], v.x.l.v[7]);
    q[2] = make_uint4(v.y.l.v[0], v.y.l.v[1], v.y.l.v[2], v.y.l.v[3]);
    q[3] = make_uint4(v.y.l.v[4], v.y.l.v[5], v.y.l.v[6], v.y.l.v[7]);
#else
    *p = v;
#endif
}

FF_HD Fq ba_load_fq(const Fq* p) {
#ifdef __CUDA_ARCH__
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fq r;
    r.l.v[0] = a.x; r.l.v[1] = a.y; r.l.v[2] = a.z; r.l.v[3] = a.w;
    r.l.v[4] = b.x; r.l.v[5] = b.y; r.l.v[6] = b.z; r.l.v[7] = b.w;
    return r;
#else
    return *p;
#endif
}
FF_HD void ba_store_fq(Fq* p, const Fq& v) {
#ifdef __CUDA_ARCH__
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l.v[0], v.l.v[1], v.l.v[2], v.l.v[3]);
    q[1] = make_uint4(v.l.v[4], v.l.v[5], v.l.v[6], v.l.v[7]);
#else
    *p = v;
#endif
}

// input `pos` of a level: a signed base at level 0, a stored partial sum above
FF_HD Affine ba_input(const BaLevel& lv, uint32_t pos) {
    if (lv.entries) {
        uint32_t e = lv.entries[pos];
        Affine p = ba_load(lv.bases + (e & 0x7fffffffu));
        if ((e & 0x80000000u) && !p.is_identity()) p.y = p.y.neg();
        return p;
    }
    return ba_load(lv.points + pos);
}

enum { BA_COPY_A = 0, BA_COPY_B = 1, BA_IDENTITY = 2, BA_ADD = 3, BA_DOUBLE = 4 };

// what output (a, b) needs and the denominator to invert for it (zero when no inversion is needed)
FF_HD int ba_classify(const Affine& a, const Affine& b, bool has_b, Fq& d) {
    d = Fq::zero();
    if (!has_b || b.is_identity()) return BA_COPY_A;
    if (a.is_identity()) return BA_COPY_B;
    if (a.x == b.x) {
        if (a.y == b.y && !a.y.is_zero()) {
            d = a.y.dbl();
            return BA_DOUBLE;
        }
        return BA_IDENTITY;  // b = -a (or a point of order two, which G1 does not have)
    }
    d = b.x - a.x;
    return BA_ADD;
}

FF_HD Affine ba_combine(int kind, const Affine& a, const Affine& b, const Fq& inv_d) {
    if (kind == BA_COPY_A) return a;
    if (kind == BA_COPY_B) return b;
    Affine r;
    if (kind == BA_IDENTITY) {
        r.x = Fq::zero();
        r.y = Fq::zero();
        return r;
    }
    Fq lambda, x3;
    if (kind == BA_ADD) {
        lambda = (b.y - a.y) * inv_d;
        x3 = lambda.sqr() - a.x - b.x;
    } else {
        Fq xx = a.x.sqr();
        lambda = (xx.dbl() + xx) * inv_d;
        x3 = lambda.sqr() - a.x.dbl();
    }
    r.x = x3;
    r.y = lambda * (a.x - x3) - a.y;
    return r;
}

// largest b with off[b] <= o and off[b + 1] > o is found by the callers' walks; this gives a starting point
FF_HD uint64_t ba_find_bucket(const uint32_t* off, uint64_t NB, uint32_t o) {
    uint64_t lo = 0, hi = NB;  // largest b with off[b] <= o
    while (hi - lo > 1) {
        uint64_t mid = (lo + hi) >> 1;
        if (off[mid] <= o) lo = mid; else hi = mid;
    }
    return lo;
}

// pass A of thread t: prefix products of the denominators of outputs [t*L, min((t+1)*L, M_out))
FF_HD void ba_thread_a(uint64_t t, uint32_t L, const BaLevel& lv, Fq* prefix, Fq* totals) {
    const uint32_t M = lv.off_out[lv.NB];
    uint64_t start = t * (uint64_t)L;
    if (start >= M) {
        ba_store_fq(totals + t, Fq::one());  // the launch covers an upper bound of M: idle threads must not poison the shared inversion
        return;
    }
    uint32_t end = (start + L < M) ? (uint32_t)(start + L) : M;
    uint64_t b = ba_find_bucket(lv.off_out, lv.NB, (uint32_t)start);
    Fq acc = Fq::one();
    for (uint32_t o = (uint32_t)start; o < end; ++o) {
        while (o >= lv.off_out[b + 1]) ++b;  // skip finished / empty buckets
        uint32_t j = o - lv.off_out[b], m_in = lv.off_in[b + 1] - lv.off_in[b], i0 = lv.off_in[b] + 2 * j;
        bool has_b = 2 * j + 1 < m_in;
        Affine pa = ba_input(lv, i0), pb = pa;
        if (has_b) pb = ba_input(lv, i0 + 1);
        Fq d;
        ba_classify(pa, pb, has_b, d);
        ba_store_fq(prefix + o, acc);
        if (!d.is_zero()) acc = acc * d;
    }
    ba_store_fq(totals + t, acc);
}

// pass B of thread t: inv_totals[t] = 1 / totals[t]; walks the same slots backwards and writes the sums
FF_HD void ba_thread_b(uint64_t t, uint32_t L, const BaLevel& lv, const Fq* prefix, const Fq* inv_totals, Affine* out) {
    const uint32_t M = lv.off_out[lv.NB];
    uint64_t start = t * (uint64_t)L;
    if (start >= M) return;
    uint32_t end = (start + L < M) ? (uint32_t)(start + L) : M;
    uint64_t b = ba_find_bucket(lv.off_out, lv.NB, end - 1);
    Fq acc = ba_load_fq(inv_totals + t);
    for (uint32_t o = end; o-- > (uint32_t)start;) {
        while (o < lv.off_out[b]) --b;
        uint32_t j = o - lv.off_out[b], m_in = lv.off_in[b + 1] - lv.off_in[b], i0 = lv.off_in[b] + 2 * j;
        bool has_b = 2 * j + 1 < m_in;
        Affine pa = ba_input(lv, i0), pb = pa;
        if (has_b) pb = ba_input(lv, i0 + 1);
        Fq d, inv_d = Fq::zero();
        int kind = ba_classify(pa, pb, has_b, d);
        if (!d.is_zero()) {
            inv_d = ba_load_fq(prefix + o) * acc;  // (d_0 .. d_{o-1}) * (d_0 .. d_o)^-1
            acc = acc * d;
        }
        ba_store(out + o, ba_combine(kind, pa, pb, inv_d));
    }
}

// bucket b of the final level (every bucket holds at most one point) -> the XYZZ bucket array the reduction reads
FF_HD XYZZ ba_final_bucket(const BaLevel& lv, uint64_t b) {
    uint32_t m = lv.off_in[b + 1] - lv.off_in[b];
    if (m == 0) return XYZZ::identity();
    return xyzz_from_affine(ba_input(lv, lv.off_in[b]));
}

}  // namespace b200zk
