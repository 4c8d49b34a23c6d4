// Host-side compiler of GraphEvaluator programs (plain C++, no CUDA) -- used by quotient.cu and, for CPU tests of the
// compile step and of the interpreter, by tests/host_emul.
//
// Upstream (halo2_proofs/src/plonk/evaluation.rs @ e5ddf67, pin /root/reference/Cargo.lock:1886-1888) keeps one
// intermediate per calculation in a Vec<F> per thread.  On the device the live values of a row sit in shared memory,
// so the program is lowered first: dead calculations are dropped; the rest is re-emitted demand-driven from the result
// (a value is computed right before its first reader, so `value = value * y + gate_i` chains keep ONE gate value live,
// not all of them); slots are recycled by outstanding-use counts; Horner becomes (MOV +) a run of MADs; every
// ValueSource becomes one operand word.  Fewer live slots = more rows resident per SM.
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
struct alignas(16) GInstr {  // one 128-bit load per instruction
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
    // liveness: only what the last calculation (the result) depends on is evaluated; uses[j] counts the operand
    // occurrences that still have to read intermediate j
    std::vector<char> needed(n_calcs, 0);
    std::vector<uint32_t> uses(n_calcs, 0);
    needed[n_calcs - 1] = 1;
    for (uint32_t i = n_calcs; i-- > 0;) {
        if (!needed[i]) continue;
        for (const auto& s : src[i])
            if (s.kind == B200ZK_SRC_INTERMEDIATE) {
                needed[s.index] = 1;
                uses[s.index]++;
            }
    }
    uses[n_calcs - 1]++;  // the result outlives the program
    std::vector<uint32_t> slot_of(n_calcs, 0), free_slots;
    std::vector<char> done(n_calcs, 0);
    uint32_t next_slot = G_SLOT_FIRST;
    bool overflow = false;
    auto alloc = [&]() -> uint32_t {
        if (!free_slots.empty()) {
            uint32_t s = free_slots.back();
            free_slots.pop_back();
            return s;
        }
        if (next_slot >= G_MAX_SLOTS) {
            overflow = true;
            return G_MAX_SLOTS - 1;
        }
        return next_slot++;
    };
    auto consume = [&](const b200zk_value_source& s) {  // one operand occurrence has been read
        if (s.kind == B200ZK_SRC_INTERMEDIATE && --uses[s.index] == 0) free_slots.push_back(slot_of[s.index]);
    };
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
    auto emit = [&](uint32_t op, uint32_t dst, uint32_t a, uint32_t b) { P.instrs.push_back(GInstr{op | (dst << 8), a, b, 0}); };
    // Demand-driven emission from the result (explicit stack: programs can be deep): an intermediate is computed right
    // before its first reader, so a Horner over many gate values keeps one of them live at a time instead of all.
    // A frame walks the operands of one calculation; `cur` is the next operand to make available, and for a Horner
    // the MOV / MADs are emitted as the walk passes the start value / each part.
    struct Frame {
        uint32_t i, cur, dst;
    };
    std::vector<Frame> stack;
    stack.push_back(Frame{n_calcs - 1, 0, 0});
    while (!stack.empty()) {
        Frame& f = stack.back();
        const uint32_t i = f.i;
        const b200zk_calculation& c = calcs[i];
        const auto& sv = src[i];
        const bool horner = c.op == B200ZK_CALC_HORNER;
        bool descended = false;
        while (f.cur < sv.size()) {
            const b200zk_value_source& s = sv[f.cur];
            if (s.kind == B200ZK_SRC_INTERMEDIATE && !done[s.index]) {
                stack.push_back(Frame{s.index, 0, 0});  // invalidates f
                descended = true;
                break;
            }
            if (horner) {
                // sv = [start, factor, parts...]: the factor must be ready before the first MAD, so operand 1 is only
                // waited for here; emission happens when the walk passes operand 0 (after 1 is ready) and each part
                if (f.cur == 1) {
                    // start and factor are both available: open the accumulator.  It may take over the start value's
                    // slot when this is that value's last reader (no MOV); otherwise a fresh slot -- never one that a
                    // pending part or the factor still occupies, because those still have a use outstanding.
                    if (c.a.kind == B200ZK_SRC_INTERMEDIATE && uses[c.a.index] == 1) {
                        f.dst = slot_of[c.a.index];
                        uses[c.a.index] = 0;
                    } else {
                        uint32_t a = encode(c.a);
                        f.dst = alloc();
                        emit(GI_MOV, f.dst, a, 0);
                        consume(c.a);
                    }
                } else if (f.cur >= 2) {
                    emit(GI_MAD, f.dst, encode(s), encode(c.b));
                    consume(s);
                }
            }
            f.cur++;
        }
        if (descended) continue;
        if (horner) {
            consume(c.b);
        } else {
            // operands are read before the destination is written, so the destination may reuse an operand's slot
            uint32_t a = encode(c.a), b = (sv.size() > 1) ? encode(c.b) : 0;
            consume(c.a);
            if (sv.size() > 1) consume(c.b);
            f.dst = alloc();
            switch (c.op) {
                case B200ZK_CALC_ADD: emit(GI_ADD, f.dst, a, b); break;
                case B200ZK_CALC_SUB: emit(GI_SUB, f.dst, a, b); break;
                case B200ZK_CALC_MUL: emit(GI_MUL, f.dst, a, b); break;
                case B200ZK_CALC_SQUARE: emit(GI_SQR, f.dst, a, 0); break;
                case B200ZK_CALC_DOUBLE: emit(GI_DBL, f.dst, a, 0); break;
                case B200ZK_CALC_NEGATE: emit(GI_NEG, f.dst, a, 0); break;
                default: emit(GI_MOV, f.dst, a, 0); break;
            }
        }
        slot_of[i] = f.dst;
        done[i] = 1;
        stack.pop_back();
        if (overflow) return "program keeps too many intermediates live at once (split it with PreviousValue)";
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
