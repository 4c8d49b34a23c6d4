// Host-emulation harness: compiles the product's device field header (ff.cuh) for the CPU, where the
// PTX carry-chain leaves are replaced by their 64-bit emulation, so the limb choreography of the
// Montgomery multiplier can be checked in the GPU-less container (tests/test_ff_host_emul.py).
#include "../../scroll-prover_b200/csrc/ff.cuh"
#include <cstring>
using namespace b200zk;
template <class F> static void binop(int op, uint32_t* r, const uint32_t* a, const uint32_t* b) {
    F x, y, z;
    memcpy(x.l.v, a, 32);
    memcpy(y.l.v, b, 32);
    switch (op) {
        case 0: z = x * y; break;
        case 1: z = x + y; break;
        case 2: z = x - y; break;
        case 3: z = x.inv(); break;
        case 4: z = x.from_mont(); break;
        case 5: z = x.to_mont(); break;
        case 6: z = x.neg(); break;
        default: z = x.sqr_scan(); break;
    }
    memcpy(r, z.l.v, 32);
}
extern "C" void ff_host_op(int field, int op, uint32_t* r, const uint32_t* a, const uint32_t* b, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        if (field == 0) binop<Fr>(op, r + 8 * i, a + 8 * i, b + 8 * i);
        else binop<Fq>(op, r + 8 * i, a + 8 * i, b + 8 * i);
    }
}

// ---- ec.cuh under host emulation: acc (Jacobian in) op q -> affine out
#include "../../scroll-prover_b200/csrc/ec.cuh"
// op 0: acc += q_affine (xyzz_madd)   1: acc += q (as xyzz from affine, via xyzz_add after making Z non-trivial)   2: 2*acc
extern "C" void ec_host_op(int op, uint32_t* out_affine16, const uint32_t* acc_jac24, const uint32_t* q_affine16) {
    Jacobian j;
    memcpy(&j, acc_jac24, 96);
    Affine q;
    memcpy(&q, q_affine16, 64);
    XYZZ acc = xyzz_from_jacobian(j);
    if (op == 0) {
        if (!q.is_identity()) xyzz_madd(acc, q.x, q.y);
    } else if (op == 1) {
        XYZZ qq = xyzz_from_affine(q);
        if (!qq.is_identity()) {  // give q a non-trivial ZZ/ZZZ: (X l^2, Y l^3, ZZ l^2, ZZZ l^3) with l = 3
            Fq l = Fq::one() + Fq::one() + Fq::one();
            Fq l2 = l.sqr(), l3 = l2 * l;
            qq.x = qq.x * l2; qq.y = qq.y * l3; qq.zz = qq.zz * l2; qq.zzz = qq.zzz * l3;
        }
        xyzz_add(acc, qq);
    } else {
        acc = xyzz_dbl(acc);
    }
    Jacobian r = xyzz_to_jacobian_normalized(acc);
    Affine a;
    a.x = r.z.is_zero() ? Fq::zero() : r.x;
    a.y = r.z.is_zero() ? Fq::zero() : r.y;
    memcpy(out_affine16, &a, 64);
}
