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

// ---- graph.hpp + graph_exec.cuh under host emulation: the product's program lowering and row interpreter run on
// the CPU with plain-array slots (tests/test_graph_host_emul.py); only the CUDA kernel wrapper is not exercised here.
#include "../../scroll-prover_b200/csrc/graph_exec.cuh"
#include <vector>
namespace {
struct ArrSlots {
    std::vector<Fr>* v;
    Fr load(uint32_t s) const { return (*v)[s]; }
    void store(uint32_t s, const Fr& x) { (*v)[s] = x; }
};
struct ArrCols {
    const Fr* const* cols;
    const uint32_t* rot_off;
    uint64_t row, mask;
    Fr load(uint32_t col, uint32_t rot) const { return cols[col][(row + rot_off[rot]) & mask]; }
};
struct ArrConsts {
    const Fr* c;
    Fr load(uint32_t i) const { return c[i]; }
};
}  // namespace
// returns 0 and fills info[0] = instructions, info[1] = slots; or -1 with the compile error in err (<= 255 chars)
extern "C" int graph_host_eval(const b200zk_calculation* calcs, uint32_t n_calcs, const b200zk_value_source* parts, uint32_t n_parts,
                               const uint32_t* constants, uint32_t n_constants, const int32_t* rotations, uint32_t n_rotations,
                               const void* const* fixed, uint32_t n_fixed, const void* const* advice, uint32_t n_advice,
                               const void* const* instance, uint32_t n_instance, const uint32_t* challenges, uint32_t n_challenges,
                               const uint32_t* bgty /* beta gamma theta y */, const uint32_t* ext_omega, uint32_t* values,
                               uint32_t log_size, int32_t rot_scale, uint32_t* info, char* err) {
    GraphProgram P;
    std::string e = graph_compile(calcs, n_calcs, parts, n_parts, n_constants, n_rotations, &P);
    if (e.empty() && (P.need_cols[0] > n_fixed || P.need_cols[1] > n_advice || P.need_cols[2] > n_instance ||
                      P.need_challenges > n_challenges))
        e = "not enough columns / challenges";
    if (!e.empty()) {
        snprintf(err, 256, "%s", e.c_str());
        return -1;
    }
    info[0] = (uint32_t)P.instrs.size();
    info[1] = P.n_slots;
    std::vector<Fr> consts(n_constants + 4 + P.need_challenges);
    if (n_constants) memcpy(consts.data(), constants, 32 * (size_t)n_constants);
    memcpy(&consts[n_constants], bgty, 128);
    if (P.need_challenges) memcpy(&consts[n_constants + 4], challenges, 32 * (size_t)P.need_challenges);
    std::vector<const Fr*> cols;
    for (uint32_t i = 0; i < P.need_cols[0]; ++i) cols.push_back((const Fr*)fixed[i]);
    for (uint32_t i = 0; i < P.need_cols[1]; ++i) cols.push_back((const Fr*)advice[i]);
    for (uint32_t i = 0; i < P.need_cols[2]; ++i) cols.push_back((const Fr*)instance[i]);
    cols.push_back(nullptr);
    const uint64_t size = 1ull << log_size;
    std::vector<uint32_t> rot_off(n_rotations + 1);
    for (uint32_t r = 0; r < n_rotations; ++r) {
        int64_t v = ((int64_t)rotations[r] * rot_scale) % (int64_t)size;
        if (v < 0) v += (int64_t)size;
        rot_off[r] = (uint32_t)v;
    }
    Fr zeta, w = Fr::one(), x;
    {   // Fr::ZETA (same constant as ntt.cu's host_zeta)
        const uint32_t v[8] = {0x55fcd653u, 0x0363f299u, 0x5fc1e200u, 0x73e7950bu, 0x576d9d24u, 0xc5fce83eu, 0xa1c3a4d4u, 0x059c805du};
        memcpy(zeta.l.v, v, 32);
    }
    Fr om = Fr::one();
    if (ext_omega) memcpy(om.l.v, ext_omega, 32);
    std::vector<Fr> slots(P.n_slots);
    Fr* vals = (Fr*)values;
    for (uint64_t row = 0; row < size; ++row) {
        ArrSlots S{&slots};
        if (P.uses_prev) S.store(G_SLOT_PREV, vals[row]);
        if (P.uses_x) S.store(G_SLOT_X, zeta * w);
        ArrCols Cc{cols.data(), rot_off.data(), row, size - 1};
        ArrConsts K{consts.data()};
        graph_exec_row(P.instrs.data(), (uint32_t)P.instrs.size(), S, Cc, K);
        vals[row] = (P.out_slot == G_NO_RESULT) ? Fr::zero() : S.load(P.out_slot);
        w = w * om;
    }
    return 0;
}
