// Witness-graph VM node semantics (circom-witnesscalc graph), one proof per thread.
// Follows rln/src/circuit/iden3calc/graph.rs:71-143 (Operation::eval_fr), :174-222 (Neg, TernCond),
// :314-466 (shl/shr/bit ops/signed comparisons).  Values are Fr in Montgomery form; the integer
// view needed by shifts, bit-ops and comparisons is obtained with one Montgomery reduction.
#pragma once
#include "fp.cuh"

namespace zk {

enum VmKind : u32 { VM_INPUT = 0, VM_CONST = 1, VM_UNO = 2, VM_DUO = 3, VM_TRES = 4 };
// VM_TRES with op 0 is the graph's TernCond; op 1 is a·b + c, which only the depth-reducing rewrite of host_util.hpp produces
constexpr u32 VM_TRES_FMA = 1;
// DuoOp numbering of the protobuf schema (rln/src/circuit/iden3calc/proto.rs:84-106)
enum VmDuo : u32 {
    OP_MUL = 0, OP_DIV, OP_ADD, OP_SUB, OP_POW, OP_IDIV, OP_MOD, OP_EQ, OP_NEQ, OP_LT, OP_GT, OP_LEQ, OP_GEQ,
    OP_LAND, OP_LOR, OP_SHL, OP_SHR, OP_BOR, OP_BAND, OP_BXOR
};

// bundle schedule of the graph (host_util.hpp vm_build_schedule, k_prover.cu k_witness)
constexpr u32 VM_SLOTS = 8;    // nodes per bundle = warps per CTA (two per scheduler: they are latency-bound, so they interleave for free)
constexpr u32 VM_RING = 8;     // bundles whose values stay in the shared-memory ring
enum VmSrc : u32 { VM_SRC_RING = 0, VM_SRC_CONST = 1, VM_SRC_GLOBAL = 2 };

struct VmInstr {  // 16 bytes: one 128-bit load per node
    u32 kind_op;  // kind | op << 8
    u32 a, b, c;
};

HD bool vm_raw_gt_half(const u32* x) {  // x > (r−1)/2 : "negative" in circom's signed view
    const u32 h[8] = {0xf8000000u, 0xa1f0fac9u, 0x3cdcb848u, 0x9419f424u, 0x40c0ac2eu, 0xdc2822dbu, 0x7098d014u, 0x18322739u};
    return Fr::raw_cmp(x, h) > 0;
}

// 256-bit integer division for Idiv / Mod (bit-serial; these ops do not occur in the RLN graphs)
HDN void vm_divmod(const u32* n, const u32* d, u32* q, u32* rem) {
    for (int i = 0; i < 8; i++) { q[i] = 0; rem[i] = 0; }
    for (int i = 255; i >= 0; i--) {
        u32 top = rem[7] >> 31;
        for (int k = 7; k > 0; k--) rem[k] = (rem[k] << 1) | (rem[k - 1] >> 31);
        rem[0] = (rem[0] << 1) | ((n[i >> 5] >> (i & 31)) & 1);
        if (top || Fr::raw_cmp(rem, d) >= 0) {
            u32 t[8];
            Fr::raw_sub(t, rem, d);
            for (int k = 0; k < 8; k++) rem[k] = t[k];
            q[i >> 5] |= 1u << (i & 31);
        }
    }
}

// returns false where the reference returns Err (value does not fit the field)
HDN bool vm_eval_duo(u32 op, const Fr& a, const Fr& b, Fr& out) {
    switch (op) {
        case OP_MUL: out = a * b; return true;
        case OP_ADD: out = a + b; return true;
        case OP_SUB: out = a - b; return true;
        case OP_DIV: out = b.is_zero() ? Fr::zero() : a * b.inv(); return true;
        case OP_EQ: out = (a == b) ? Fr::one() : Fr::zero(); return true;
        case OP_NEQ: out = (a == b) ? Fr::zero() : Fr::one(); return true;
        case OP_LAND: out = (a.is_zero() || b.is_zero()) ? Fr::zero() : Fr::one(); return true;
        case OP_LOR: out = (a.is_zero() && b.is_zero()) ? Fr::zero() : Fr::one(); return true;
        default: break;
    }
    u32 x[8], y[8], r[8], p[8];
    a.to_canonical(x);
    b.to_canonical(y);
    for (int i = 0; i < 8; i++) p[i] = FrCfg::p(i);
    switch (op) {
        case OP_POW: out = a.pow(y); return true;
        case OP_IDIV:
        case OP_MOD: {
            bool bz = true;
            for (int i = 0; i < 8; i++) bz = bz && y[i] == 0;
            if (bz) { out = Fr::zero(); return true; }
            u32 q[8], rem[8];
            vm_divmod(x, y, q, rem);
            out = Fr::from_canonical(op == OP_IDIV ? q : rem);
            return true;
        }
        case OP_LT: case OP_GT: case OP_LEQ: case OP_GEQ: {
            bool xn = vm_raw_gt_half(x), yn = vm_raw_gt_half(y);
            int c = Fr::raw_cmp(x, y);
            bool res;
            if (xn != yn) { bool lt = xn; res = (op == OP_LT || op == OP_LEQ) ? lt : !lt; }
            else res = op == OP_LT ? c < 0 : op == OP_GT ? c > 0 : op == OP_LEQ ? c <= 0 : c >= 0;
            out = res ? Fr::one() : Fr::zero();
            return true;
        }
        case OP_SHL: case OP_SHR: {
            bool bz = true, small = true;
            for (int i = 0; i < 8; i++) bz = bz && y[i] == 0;
            for (int i = 1; i < 8; i++) small = small && y[i] == 0;
            if (bz) { out = a; return true; }
            if (!small || y[0] >= 254) { out = Fr::zero(); return true; }
            const u32 s = y[0], ws = s >> 5, bs = s & 31;
            for (int i = 0; i < 8; i++) {
                u32 v = 0;
                if (op == OP_SHR) {
                    u32 src = i + ws;
                    if (src < 8) v = x[src] >> bs;
                    if (bs && src + 1 < 8) v |= x[src + 1] << (32 - bs);
                } else {
                    int src = i - (int)ws;
                    if (src >= 0) v = x[src] << bs;
                    if (bs && src - 1 >= 0) v |= x[src - 1] >> (32 - bs);
                }
                r[i] = v;
            }
            if (Fr::raw_cmp(r, p) >= 0) return false;  // from_bigint fails (only reachable for Shl)
            out = Fr::from_canonical(r);
            return true;
        }
        case OP_BOR: case OP_BAND: case OP_BXOR: {
            for (int i = 0; i < 8; i++) r[i] = op == OP_BOR ? (x[i] | y[i]) : op == OP_BAND ? (x[i] & y[i]) : (x[i] ^ y[i]);
            if (Fr::raw_cmp(r, p) > 0) Fr::raw_sub(r, r, p);
            if (Fr::raw_cmp(r, p) >= 0) return false;
            out = Fr::from_canonical(r);
            return true;
        }
        default: return false;
    }
}

}  // namespace zk
