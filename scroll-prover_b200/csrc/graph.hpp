// Host-side compiler of GraphEvaluator programs (plain C++, no CUDA) -- used by quotient.cu and, for CPU tests of the
// compile step and of the interpreter, by tests/host_emul.
//
// Upstream (halo2_proofs/src/plonk/evaluation.rs @ e5ddf67, pin /root/reference/Cargo.lock:1886-1888) keeps one
// intermediate per calculation in a Vec<F> per thread.  On the device the live values of a row sit in shared memory,
// so the program is lowered first: dead calculations are dropped, every surviving one gets an on-chip slot by a
// linear scan over last uses, Horner becomes MOV + a run of MADs, and every ValueSource becomes one operand word.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/b200zk.h"

namespace b200zk {

enum { GI_ADD = 0, GI_SUB, GI_MUL, GI_SQR, GI_DBL, GI_NEG, GI_MOV, GI_MAD };  // MAD: dst = dst * b + a
enum { GK_CONST = 0, GK_SLOT = 1, GK_COL = 2 };
// operand word = kind << 28 | payload
//   GK_CONST payload: index into [program constants | beta gamma theta y | challenges...]
//   GK_SLOT  payload: slot (0 = PreviousValue, 1 = extended X, >= 2 intermediates)
//   GK_COL   payload: rotation index << 16 | position in the call's column table, which lists the columns the program
//            reads as fixed[0..need_cols[0]) | advice[0..need_cols[1]) | instance[0..need_cols[2])
struct GInstr {
    uint32_t op_dst;  // op | dst slot << 8
    uint32_t a, b, pad;
};
constexpr uint32_t G_SLOT_PREV = 0, G_SLOT_X = 1, G_SLOT_FIRST = 2;
constexpr uint32_t G_MAX_SLOTS = 224;     // 32 rows per block x 224 slots x 32 B = 224 KB of shared memory
constexpr uint32_t G_MAX_ROTATIONS = 1024, G_MAX_COLUMNS = 65536;
constexpr uint32_t G_NO_RESULT = 0xffffffffu;

struct GraphProgram {
    std::vector<GInstr> instrs;
    uint32_t n_slots = G_SLOT_FIRST;  // reserved slots included
    uint32_t out_slot = G_NO_RESULT;  // G_NO_RESULT: no calculations, GraphEvaluator::evaluate returns zero
    uint32_t n_constants = 0, n_rotations = 0;
    uint32_t need_cols[3] = {0, 0, 0};  // 1 + highest column index read per table
    uint32_t need_challenges = 0;       // 1 + highest challenge index read
    bool uses_x = false, uses_prev = false;
};

inline uint32_t g_operand(uint32_t kind, uint32_t payload) { return (kind << 28) | payload; }

// returns "" on success, else the reason the program is rejected
inline std::string graph_compile(const b200zk_calculation* calcs, uint32_t n_calcs, const b200zk_value_source* parts,
                                 uint32_t n_parts, uint32_t n_constants, uint32_t n_rotations, GraphProgram* out) {
    GraphProgram& P = *out;
    P = GraphProgram();
    P.n_constants = n_constants;
    P.n_rotations = n_rotations;
    if (n_rotations > G_MAX_ROTATIONS) return "too many rotations";
    if (n_constants >= (1u << 24)) return "too many constants";
    char msg[160];
    // every ValueSource a calculation reads, in evaluation order
    auto sources_of = [&](uint32_t i, std::vector<b200zk_value_source>& v) -> const char* {
        const b200zk_calculation& c = calcs[i];
        v.clear();
        switch (c.op) {
            case B200ZK_CALC_ADD: case B200ZK_CALC_SUB: case B200ZK_CALC_MUL:
                v.push_back(c.a); v.push_back(c.b); break;
            case B200ZK_CALC_SQUARE: case B200ZK_CALC_DOUBLE: case B200ZK_CALC_NEGATE: case B200ZK_CALC_STORE:
                v.push_back(c.a); break;
            case B200ZK_CALC_HORNER:
                if ((uint64_t)c.parts_offset + c.parts_len > n_parts) return "Horner parts out of range";
                v.push_back(c.a);
                v.push_back(c.b);
                for (uint32_t j = 0; j < c.parts_len; ++j) v.push_back(parts[c.parts_offset + j]);
                break;
            default: return "unknown calculation";
        }
        return nullptr;
    };
    std::vector<std::vector<b200zk_value_source>> src(n_calcs);
    for (uint32_t i = 0; i < n_calcs; ++i) {
        if (const char* e = sources_of(i, src[i])) {
            snprintf(msg, sizeof msg, "calculation %u: %s", i, e);
            return msg;
        }
        for (const auto& s : src[i]) {
            const char* e = nullptr;
            switch (s.kind) {
                case B200ZK_SRC_CONSTANT: if (s.index >= n_constants) e = "constant index out of range"; break;
                case B200ZK_SRC_INTERMEDIATE: if (s.index >= i) e = "intermediate is not an earlier calculation"; break;
                case B200ZK_SRC_FIXED: case B200ZK_SRC_ADVICE: case B200ZK_SRC_INSTANCE:
                    if (s.index >= G_MAX_COLUMNS) e = "column index out of range";
                    else if (s.rotation >= n_rotations) e = "rotation index out of range";
                    break;
                case B200ZK_SRC_CHALLENGE: if (s.index >= (1u << 20)) e = "challenge index out of range"; break;
                case B200ZK_SRC_BETA: case B200ZK_SRC_GAMMA: case B200ZK_SRC_THETA: case B200ZK_SRC_Y:
                case B200ZK_SRC_PREVIOUS_VALUE: case B200ZK_SRC_EXTENDED_X: break;
                default: e = "unknown value source";
            }
            if (e) {
                snprintf(msg, sizeof msg, "calculation %u: %s", i, e);
                return msg;
            }
        }
    }
    if (n_calcs == 0) return "";
    // liveness: only what the last calculation (the result) depends on is evaluated
    std::vector<char> needed(n_calcs, 0);
    std::vector<uint32_t> last_use(n_calcs, 0);
    needed[n_calcs - 1] = 1;
    last_use[n_calcs - 1] = n_calcs;  // the result outlives the program
    for (uint32_t i = n_calcs; i-- > 0;) {
        if (!needed[i]) continue;
        for (const auto& s : src[i])
            if (s.kind == B200ZK_SRC_INTERMEDIATE) {
                if (!needed[s.index]) {
                    needed[s.index] = 1;
                    last_use[s.index] = i;  // i descends: the first visit is the last use
                }
            }
    }
    std::vector<uint32_t> slot_of(n_calcs, 0), free_slots;
    uint32_t next_slot = G_SLOT_FIRST;
    auto encode = [&](const b200zk_value_source& s) -> uint32_t {
        switch (s.kind) {
            case B200ZK_SRC_CONSTANT: return g_operand(GK_CONST, s.index);
            case B200ZK_SRC_INTERMEDIATE: return g_operand(GK_SLOT, slot_of[s.index]);
            case B200ZK_SRC_FIXED: case B200ZK_SRC_ADVICE: case B200ZK_SRC_INSTANCE: {
                uint32_t t = s.kind - B200ZK_SRC_FIXED;
                if (s.index + 1 > P.need_cols[t]) P.need_cols[t] = s.index + 1;
                return g_operand(GK_COL, (t << 26) | (s.rotation << 16) | s.index);
            }
            case B200ZK_SRC_CHALLENGE:
                if (s.index + 1 > P.need_challenges) P.need_challenges = s.index + 1;
                return g_operand(GK_CONST, n_constants + 4 + s.index);
            case B200ZK_SRC_BETA: return g_operand(GK_CONST, n_constants + 0);
            case B200ZK_SRC_GAMMA: return g_operand(GK_CONST, n_constants + 1);
            case B200ZK_SRC_THETA: return g_operand(GK_CONST, n_constants + 2);
            case B200ZK_SRC_Y: return g_operand(GK_CONST, n_constants + 3);
            case B200ZK_SRC_PREVIOUS_VALUE: P.uses_prev = true; return g_operand(GK_SLOT, G_SLOT_PREV);
            default: P.uses_x = true; return g_operand(GK_SLOT, G_SLOT_X);
        }
    };
    for (uint32_t i = 0; i < n_calcs; ++i) {
        if (!needed[i]) continue;
        const b200zk_calculation& c = calcs[i];
        // the destination is taken before the operands are released: a MAD run rewrites dst while its parts are live
        uint32_t dst;
        if (!free_slots.empty()) {
            dst = free_slots.back();
            free_slots.pop_back();
        } else {
            dst = next_slot++;
            if (next_slot > G_MAX_SLOTS) return "program keeps too many intermediates live at once (split it with PreviousValue)";
        }
        slot_of[i] = dst;
        auto emit = [&](uint32_t op, uint32_t a, uint32_t b) { P.instrs.push_back(GInstr{op | (dst << 8), a, b, 0}); };
        switch (c.op) {
            case B200ZK_CALC_ADD: emit(GI_ADD, encode(c.a), encode(c.b)); break;
            case B200ZK_CALC_SUB: emit(GI_SUB, encode(c.a), encode(c.b)); break;
            case B200ZK_CALC_MUL: emit(GI_MUL, encode(c.a), encode(c.b)); break;
            case B200ZK_CALC_SQUARE: emit(GI_SQR, encode(c.a), 0); break;
            case B200ZK_CALC_DOUBLE: emit(GI_DBL, encode(c.a), 0); break;
            case B200ZK_CALC_NEGATE: emit(GI_NEG, encode(c.a), 0); break;
            case B200ZK_CALC_STORE: emit(GI_MOV, encode(c.a), 0); break;
            default: {  // Horner(start, parts, factor): value = start; for part: value = value * factor + part
                emit(GI_MOV, encode(c.a), 0);
                uint32_t f = encode(c.b);
                for (uint32_t j = 0; j < c.parts_len; ++j) emit(GI_MAD, encode(parts[c.parts_offset + j]), f);
            }
        }
        for (const auto& s : src[i])
            if (s.kind == B200ZK_SRC_INTERMEDIATE && last_use[s.index] == i) {
                free_slots.push_back(slot_of[s.index]);
                last_use[s.index] = 0xffffffffu;  // release once even when named twice
            }
    }
    P.n_slots = next_slot;
    P.out_slot = slot_of[n_calcs - 1];
    // column operands were emitted as (table, column); now that every table's extent is known, flatten them
    const uint32_t base[3] = {0, P.need_cols[0], P.need_cols[0] + P.need_cols[1]};
    if ((uint64_t)base[2] + P.need_cols[2] > G_MAX_COLUMNS) return "too many columns";
    auto flatten = [&](uint32_t& w) {
        if ((w >> 28) != GK_COL) return;
        uint32_t pay = w & 0x0fffffffu, t = pay >> 26, rot = (pay >> 16) & 0x3ffu, col = pay & 0xffffu;
        w = g_operand(GK_COL, (rot << 16) | (base[t] + col));
    };
    for (auto& ins : P.instrs) {
        flatten(ins.a);
        flatten(ins.b);
    }
    return "";
}

}  // namespace b200zk
