// Row interpreter of a lowered GraphEvaluator program (graph.hpp).  FF_HD: the same code runs in the CUDA kernel
// (slots in shared memory, columns in HBM) and, under host emulation, in tests/host_emul (plain arrays).
//
// Mirrors GraphEvaluator::evaluate / Calculation::evaluate (halo2_proofs/src/plonk/evaluation.rs @ e5ddf67).
#pragma once
#include "ff.cuh"
#include "graph.hpp"

namespace b200zk {

// Slots: Fr load(uint32_t slot) const; void store(uint32_t slot, const Fr&)
// Cols : Fr load(uint32_t column_table_position, uint32_t rotation_index) const
// Consts: Fr load(uint32_t index) const
template <class Slots, class Cols, class Consts>
FF_HD void graph_exec_row(const GInstr* __restrict__ instrs, uint32_t n_instr, Slots& S, const Cols& C, const Consts& K) {
    auto fetch = [&](uint32_t w) -> Fr {
        const uint32_t kind = w >> 28, pay = w & 0x0fffffffu;
        if (kind == GK_SLOT) return S.load(pay);
        if (kind == GK_CONST) return K.load(pay);
        return C.load(pay & 0xffffu, (pay >> 16) & 0x3ffu);
    };
    for (uint32_t pc = 0; pc < n_instr; ++pc) {
        const GInstr ins = instrs[pc];
        const uint32_t op = ins.op_dst & 0xffu, dst = ins.op_dst >> 8;
        Fr a = fetch(ins.a), r;
        if (op >= GI_MUL && op != GI_DBL && op != GI_NEG && op != GI_MOV) {
            // one shared multiplier: MUL a*b, SQR a*a, MAD dst*b + a
            Fr x, y;
            if (op == GI_MUL) { x = a; y = fetch(ins.b); }
            else if (op == GI_SQR) { x = a; y = a; }
            else { x = S.load(dst); y = fetch(ins.b); }
            r = x * y;
            if (op == GI_MAD) r = r + a;
        } else if (op == GI_ADD) r = a + fetch(ins.b);
        else if (op == GI_SUB) r = a - fetch(ins.b);
        else if (op == GI_DBL) r = a + a;
        else if (op == GI_NEG) r = Fr::zero() - a;
        else r = a;
        S.store(dst, r);
    }
}

}  // namespace b200zk
