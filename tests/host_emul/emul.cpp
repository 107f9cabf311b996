// TEST HARNESS (not product): compiles the product's __host__ __device__ math headers with g++ so the
// per-thread device functions can be checked on a CPU-only box against oracle/pyref before any GPU
// time is spent.  Exports a tiny C API for ctypes.  Nothing in zerokit_b200/ links this.
#include <cstring>

#include "curve.cuh"
#include "poseidon.cuh"
#include "poseidon_constants.hpp"
#include "pairing_constants.hpp"
#include "tower.cuh"
#include "vm.cuh"
#include "glv.cuh"
#include "host_util.hpp"
#include "verify_vm_program.hpp"

using namespace zk;

template <class F>
static F ld(const uint8_t* b) { u32 c[8]; memcpy(c, b, 32); return F::from_canonical(c); }
template <class F>
static void st(uint8_t* b, const F& v) { u32 c[8]; v.to_canonical(c); memcpy(b, c, 32); }

static PoseidonTables g_pt;
static bool g_pt_ready = false;
static const PoseidonTables* tables() { if (!g_pt_ready) { poseidon_fill_tables(g_pt); g_pt_ready = true; } return &g_pt; }

extern "C" {
// op: 0 mul 1 add 2 sub 3 inv 4 neg ; field: 0 Fr 1 Fq
void emu_field_op(int field, int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    if (field == 0) {
        Fr x = ld<Fr>(a), y = ld<Fr>(b), r;
        r = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : op == 3 ? x.inv() : x.neg();
        st(out, r);
    } else {
        Fq x = ld<Fq>(a), y = ld<Fq>(b), r;
        r = op == 0 ? x * y : op == 1 ? x + y : op == 2 ? x - y : op == 3 ? x.inv() : x.neg();
        st(out, r);
    }
}
void emu_poseidon(const uint8_t* in, int n, uint8_t* out) {
    Fr v[3];
    for (int i = 0; i < n; i++) v[i] = ld<Fr>(in + 32 * i);
    Fr r = n == 1 ? poseidon1(tables(), v[0]) : n == 2 ? poseidon2(tables(), v[0], v[1]) : poseidon3(tables(), v[0], v[1], v[2]);
    st(out, r);
}
// G1: k·P via XYZZ double-and-add, and P+Q via mixed add; points 64 B canonical (all-zero = infinity)
static G1Affine ld_g1(const uint8_t* b) { return {ld<Fq>(b), ld<Fq>(b + 32)}; }
static void st_g1(uint8_t* b, const G1Affine& p) { st(b, p.x); st(b + 32, p.y); }
static G2Affine ld_g2(const uint8_t* b) { return {{ld<Fq>(b), ld<Fq>(b + 32)}, {ld<Fq>(b + 64), ld<Fq>(b + 96)}}; }
static void st_g2(uint8_t* b, const G2Affine& p) { st(b, p.x.a); st(b + 32, p.x.b); st(b + 64, p.y.a); st(b + 96, p.y.b); }
void emu_g1_mul(const uint8_t* p, const uint8_t* k, uint8_t* out) {
    u32 kk[8]; memcpy(kk, k, 32);
    st_g1(out, G1XYZZ::from_affine(ld_g1(p)).mul(kk).to_affine());
}
void emu_g1_add(const uint8_t* p, const uint8_t* q, uint8_t* out) {
    G1XYZZ a = G1XYZZ::from_affine(ld_g1(p));
    G1Affine b = ld_g1(q);
    if (!b.is_inf()) a.add_affine(b);
    st_g1(out, a.to_affine());
}
void emu_g1_add_full(const uint8_t* p, const uint8_t* q, uint8_t* out) {  // XYZZ + XYZZ with non-trivial ZZ
    G1XYZZ a = G1XYZZ::from_affine(ld_g1(p)).dbl(), b = G1XYZZ::from_affine(ld_g1(q)).dbl();
    a.add(b);
    st_g1(out, a.to_affine());  // = 2P + 2Q
}
void emu_g2_mul(const uint8_t* p, const uint8_t* k, uint8_t* out) {
    u32 kk[8]; memcpy(kk, k, 32);
    st_g2(out, G2XYZZ::from_affine(ld_g2(p)).mul(kk).to_affine());
}
void emu_g2_add(const uint8_t* p, const uint8_t* q, uint8_t* out) {
    G2XYZZ a = G2XYZZ::from_affine(ld_g2(p));
    G2Affine b = ld_g2(q);
    if (!b.is_inf()) a.add_affine(b);
    st_g2(out, a.to_affine());
}
// the witness VM exactly as k_witness runs it — bundle schedule, operand sources (ring / constant table / vals), the store flag
// in bit 31 of `out` — for ONE proof: slots of a bundle are evaluated one after the other (they are independent), the ring is
// overwritten in place.  consts_resident = 1: the constant table is an operand source and dead constant nodes are skipped (the
// kernel's shared-memory mode); 0: constants are ordinary nodes.  Nodes the kernel never stores to vals stay at the poison value,
// so a consumer that reads one of them from vals shows up as a wrong wire.
// inputs: n_slots × 32 canonical bytes; out: n_nodes × 32 canonical bytes (32 × 0xff for a node that is never stored).
// returns −1 on a parse error, else the "bad" flag.
int emu_witness_scheduled(const uint8_t* graph, size_t glen, const uint8_t* inputs, uint8_t* out, uint32_t* n_bundles_out, int consts_resident,
                          uint32_t* n_stored_out) {
    GraphHost g;
    try { parse_graph(graph, glen, g); } catch (...) { return -1; }
    uint32_t nb = 0;
    std::vector<uint8_t> is_signal(g.prog.size(), 0);
    for (uint32_t node : g.signals) is_signal[node] = 1;
    std::vector<VmRecord> recs = vm_build_schedule(g.prog, nb, &is_signal, consts_resident != 0);
    if (n_bundles_out) *n_bundles_out = nb;
    Fr poison = Fr::zero();   // a value no node of these tests takes: a wrong read shows up as a wrong wire
    poison.l[0] = 0xdeadbeefu; poison.l[3] = 0x1234567u;
    std::vector<Fr> ring(VM_RING * VM_SLOTS), vals(g.prog.size(), poison), consts(g.consts.size() / 32);
    std::vector<uint8_t> is_stored(g.prog.size(), 0);
    for (size_t i = 0; i < consts.size(); i++) consts[i] = ld<Fr>(g.consts.data() + 32 * i);
    auto operand = [&](uint32_t enc) -> Fr {
        const uint32_t src = enc >> 30, idx = enc & 0x3fffffffu;
        return src == VM_SRC_RING ? ring[idx] : src == VM_SRC_CONST ? consts[idx] : vals[idx];
    };
    int bad = 0;
    uint32_t stored = 0;
    for (uint32_t b = 0; b < nb; b++) {
        Fr res[VM_SLOTS];
        bool have[VM_SLOTS];
        for (uint32_t sl = 0; sl < VM_SLOTS; sl++) {   // all slots read the ring BEFORE any slot of this bundle writes it
            const VmRecord& r = recs[(size_t)b * VM_SLOTS + sl];
            have[sl] = r.kind_op != 0xffffffffu;
            if (!have[sl]) continue;
            const uint32_t kind = r.kind_op & 0xff, op = r.kind_op >> 8;
            Fr v;
            if (kind == VM_DUO) { if (!vm_eval_duo(op, operand(r.a), operand(r.b), v)) { bad = 1; v = Fr::zero(); } }
            else if (kind == VM_CONST) v = consts[r.a];
            else if (kind == VM_INPUT) v = ld<Fr>(inputs + 32 * r.a);
            else if (kind == VM_UNO) { if (op == 0) v = operand(r.a).neg(); else { bad = 1; v = Fr::zero(); } }
            else if (op == VM_TRES_FMA) v = operand(r.a) * operand(r.b) + operand(r.c);
            else { Fr t = operand(r.a); v = t.is_zero() ? operand(r.c) : operand(r.b); }
            res[sl] = v;
        }
        for (uint32_t sl = 0; sl < VM_SLOTS; sl++) {
            if (!have[sl]) continue;
            const VmRecord& r = recs[(size_t)b * VM_SLOTS + sl];
            ring[(b % VM_RING) * VM_SLOTS + sl] = res[sl];
            if (r.out >> 31) { vals[r.out & 0x7fffffffu] = res[sl]; is_stored[r.out & 0x7fffffffu] = 1; stored++; }
        }
    }
    if (n_stored_out) *n_stored_out = stored;
    for (size_t i = 0; i < vals.size(); i++) {
        if (is_stored[i]) st(out + 32 * i, vals[i]);
        else memset(out + 32 * i, 0xff, 32);   // never stored: not a canonical value
    }
    return bad;
}
// the same for the depth-reduced program (host_util.hpp vm_optimize_program → vm_build_schedule, what the product runs): only the
// wires are comparable with the oracle.  wires_out: n_signals × 32 canonical bytes; stats: nodes, constants, bundles, stores
int emu_witness_optimized(const uint8_t* graph, size_t glen, const uint8_t* inputs, uint8_t* wires_out, uint32_t* stats) {
    GraphHost g;
    try { parse_graph(graph, glen, g); } catch (...) { return -1; }
    VmOptimized opt = vm_optimize_program(g.prog, g.consts, g.signals);
    uint32_t nb = 0;
    std::vector<uint8_t> is_signal(opt.prog.size(), 0);
    for (uint32_t node : opt.signals) is_signal[node] = 1;
    const bool resident = true;
    std::vector<VmRecord> recs = vm_build_schedule(opt.prog, nb, &is_signal, resident);
    Fr poison = Fr::zero();
    poison.l[0] = 0xdeadbeefu; poison.l[3] = 0x1234567u;
    std::vector<Fr> ring(VM_RING * VM_SLOTS), vals(opt.prog.size(), poison), consts(opt.consts.size() / 32);
    for (size_t i = 0; i < consts.size(); i++) consts[i] = ld<Fr>(opt.consts.data() + 32 * i);
    auto operand = [&](uint32_t enc) -> Fr {
        const uint32_t src = enc >> 30, idx = enc & 0x3fffffffu;
        return src == VM_SRC_RING ? ring[idx] : src == VM_SRC_CONST ? consts[idx] : vals[idx];
    };
    int bad = 0;
    uint32_t stored = 0;
    for (uint32_t b = 0; b < nb; b++) {
        Fr res[VM_SLOTS];
        bool have[VM_SLOTS];
        for (uint32_t sl = 0; sl < VM_SLOTS; sl++) {
            const VmRecord& r = recs[(size_t)b * VM_SLOTS + sl];
            have[sl] = r.kind_op != 0xffffffffu;
            if (!have[sl]) continue;
            const uint32_t kind = r.kind_op & 0xff, op = r.kind_op >> 8;
            Fr v;
            if (kind == VM_DUO) { if (!vm_eval_duo(op, operand(r.a), operand(r.b), v)) { bad = 1; v = Fr::zero(); } }
            else if (kind == VM_CONST) v = consts[r.a];
            else if (kind == VM_INPUT) v = ld<Fr>(inputs + 32 * r.a);
            else if (kind == VM_UNO) { if (op == 0) v = operand(r.a).neg(); else { bad = 1; v = Fr::zero(); } }
            else if (op == VM_TRES_FMA) v = operand(r.a) * operand(r.b) + operand(r.c);
            else { Fr t = operand(r.a); v = t.is_zero() ? operand(r.c) : operand(r.b); }
            res[sl] = v;
        }
        for (uint32_t sl = 0; sl < VM_SLOTS; sl++) {
            if (!have[sl]) continue;
            const VmRecord& r = recs[(size_t)b * VM_SLOTS + sl];
            ring[(b % VM_RING) * VM_SLOTS + sl] = res[sl];
            if (r.out >> 31) { vals[r.out & 0x7fffffffu] = res[sl]; stored++; }
        }
    }
    for (size_t w = 0; w < opt.signals.size(); w++) {
        // a wire that is a resident constant has no record: the product reads it from the constant table as well
        const VmInstr& n = opt.prog[opt.signals[w]];
        st(wires_out + 32 * w, (n.kind_op & 0xff) == VM_CONST ? consts[n.a] : vals[opt.signals[w]]);
    }
    uint32_t far = 0, operands = 0;
    for (const VmRecord& r : recs) {
        if (r.kind_op == 0xffffffffu) continue;
        const uint32_t kind = r.kind_op & 0xff;
        const uint32_t ops[3] = {r.a, r.b, r.c};
        const int no = kind == VM_UNO ? 1 : kind == VM_DUO ? 2 : kind == VM_TRES ? 3 : 0;
        for (int k = 0; k < no; k++) { operands++; if ((ops[k] >> 30) == VM_SRC_GLOBAL) far++; }
    }
    if (stats) { stats[0] = (uint32_t)opt.prog.size(); stats[1] = (uint32_t)consts.size(); stats[2] = nb; stats[3] = stored; stats[4] = far; stats[5] = operands; }
    return bad;
}
// ---- the pairing VM (verify_vm*.hpp): program built by the product's own tracer / scheduler, executed here lane by lane -------------
// One level at a time: every lane's sum is evaluated against the slots as they were before the level, lane pairs are combined,
// then the results are stored — the order the kernel's barrier enforces.  Arithmetic: the portable Montgomery product on the
// scaled / complemented operand, results summed mod q (the kernel's single-reduction sum gives the same fully reduced value).
static Fq vm_host_term(const Fq& a_in, const Fq& b, bool neg, u32 sh) {
    Fq a = neg ? a_in.neg_lazy() : a_in;
    for (u32 s = 0; s < sh; s++) { u32 t[8]; Fq::raw_add(t, a.l, a.l); memcpy(a.l, t, 32); }   // ≤ 4q < 2^256
    return a * b;
}
static int vm_host_run(const pvm::Program& P, const uint8_t* proof, const G1XYZZ& vkx, uint64_t* n_terms_out) {
    using namespace pvm;
    std::vector<Fq> slots(P.n_slots, Fq::zero());
    for (u32 i = 0; i < P.n_const; i++) slots[i] = P.consts[i];
    ProofFlags fl{0};
    u32 st = pv_prologue(proof, slots.data(), fl);
    if (st != ST_RUNNING) return (int)st;
    uint64_t n_terms = 0;
    for (u32 l = 0; l < P.n_levels; l++) {
        const u32* rec = P.code.data() + (size_t)l * REC_WORDS * LANES;
        const u32 special = (rec[1 * LANES] >> 8) & 0xff;
        if (special == SP_EXP) {
            for (int gl = 0; gl < EXP_LANES; gl++) {
                const u32 w0 = rec[gl], t0 = rec[2 * LANES + gl];
                if (!(w0 & W0_STORE)) continue;
                slots[w0 & 0xfff] = pv_pow(slots[t0 & 0xfff], P.exps[(rec[1 * LANES] >> 16) & 3], slots.data() + ((t0 >> 12) & 0xfff));
                n_terms += 330;
            }
            continue;
        }
        if (special) {
            u32 args[32];
            for (int k = 0; k < 32; k++) args[k] = rec[k];
            if (special == SP_VKX) {
                if (vkx.is_inf()) return ST_FALLBACK;
                slots[S_VX] = vkx.X; slots[S_VY] = vkx.Y; slots[S_VZZ] = vkx.ZZ; slots[S_VZZZ] = vkx.ZZZ;
            } else if (special == SP_SELECT) {
                st = pv_select(slots.data(), args, fl);
                if (st != ST_RUNNING) return (int)st;
            } else if (special == SP_FINAL) {
                if (n_terms_out) *n_terms_out = n_terms;
                return (int)pv_final(slots.data(), args);
            } else return -2;
            continue;
        }
        Fq res[LANES];
        for (int gl = 0; gl < LANES; gl++) {
            const u32 w1 = rec[1 * LANES + gl];
            const u32 N = w1 & 15;
            Fq r = Fq::zero();
            for (u32 t = 0; t < N; t++) {
                const u32 tw = rec[(2 + t) * LANES + gl];
                r = r + vm_host_term(slots[tw & 0xfff], slots[(tw >> 12) & 0xfff], (tw >> 24) & 1, (tw >> 25) & 3);
                if ((tw & 0xffffff) != 0) n_terms++;
            }
            res[gl] = r;
        }
        Fq step1[LANES];   // the two shuffle steps of the kernel: lane ^ 16, then lane ^ 8
        for (int gl = 0; gl < LANES; gl++) step1[gl] = (rec[gl] & W0_COMBINE) ? res[gl] + res[gl ^ 16] : res[gl];
        for (int gl = 0; gl < LANES; gl++) {
            const u32 w0 = rec[gl];
            if (!(w0 & W0_STORE)) continue;
            slots[w0 & 0xfff] = (w0 & W0_COMBINE4) ? step1[gl] + step1[gl ^ 8] : step1[gl];
        }
    }
    return -3;   // no SP_FINAL reached
}
// vk: alpha_g1 (64 B) | beta_g2 (128) | gamma_g2 (128) | delta_g2 (128), canonical affine; gamma_abc: (n_public + 1) × 64 B;
// publics: n_proofs × n_public × 32 B canonical scalars; proofs: n_proofs × 128 B ark-compressed.  out[j] = status of proof j
// (pvm::Status); info = {levels, slots, constants, nodes, estimated cycles, terms executed per proof}
int emu_verify_vm(const uint8_t* vk, const uint8_t* gamma_abc, int n_public, const uint8_t* publics, const uint8_t* proofs, int n_proofs, int* out,
                  uint64_t* info) {
    static PairingTables pt; static bool init = false;
    if (!init) { pairing_tables_init(pt); init = true; }
    static pvm::Program P;
    static std::vector<uint8_t> key;
    try {
        if (key.size() != 448 || memcmp(key.data(), vk, 448) != 0) {
            static FixedLines lg, ld_;
            precompute_lines(&pt, ld_g2(vk + 192), lg);
            precompute_lines(&pt, ld_g2(vk + 320), ld_);
            pvm::VerifyKeyHost h{&lg, &ld_, miller_loop(&pt, ld_g2(vk + 64), ld_g1(vk))};
            P = pvm::build_verify_program(pt, h);
            key.assign(vk, vk + 448);
        }
    } catch (const std::exception& e) {
        fprintf(stderr, "emu_verify_vm: %s\n", e.what());
        return -1;
    }
    uint64_t n_terms = 0;
    for (int j = 0; j < n_proofs; j++) {
        G1XYZZ vkx = G1XYZZ::from_affine(ld_g1(gamma_abc));
        for (int i = 0; i < n_public; i++) {
            u32 k[8];
            memcpy(k, publics + ((size_t)j * n_public + i) * 32, 32);
            vkx.add(G1XYZZ::from_affine(ld_g1(gamma_abc + 64 * (i + 1))).mul(k));
        }
        out[j] = vm_host_run(P, proofs + 128 * (size_t)j, vkx, &n_terms);
    }
    if (info) { info[0] = P.n_levels; info[1] = P.n_slots; info[2] = P.n_const; info[3] = P.n_nodes; info[4] = (uint64_t)P.est_cycles; info[5] = n_terms; }
    return 0;
}
// GLV split of a canonical scalar: out = |k1| (16 B) | |k2| (16 B) | sign1 | sign2
void emu_glv_split(const uint8_t* k, uint8_t* out) {
    u32 kk[8], h[8];
    memcpy(kk, k, 32);
    for (int half = 0; half < 2; half++) {
        out[32 + half] = glv::split(kk, half, h) ? 1 : 0;
        memcpy(out + 16 * half, h, 16);
    }
}
// kp·P + kq·Q through the Straus / GLV routine of the proof assembly
void emu_glv_double_mul(const uint8_t* p, const uint8_t* kp, const uint8_t* q, const uint8_t* kq, int use_q, uint8_t* out) {
    u32 a[8], b[8];
    memcpy(a, kp, 32);
    memcpy(b, kq, 32);
    G1XYZZ P = G1XYZZ::from_affine(ld_g1(p)), Q = G1XYZZ::from_affine(ld_g1(q));
    P = P.dbl(); P.add(G1XYZZ::from_affine(ld_g1(p)).neg());   // same point, non-trivial ZZ / ZZZ
    st_g1(out, glv_double_mul(P, a, Q, b, use_q != 0).to_affine());
}
// G2 membership test used by the verifier's point decompression (point must be on the twist)
int emu_g2_in_subgroup(const uint8_t* p) {
    static PairingTables pt; static bool init = false;
    if (!init) { pairing_tables_init(pt); init = true; }
    return g2_in_subgroup(&pt, ld_g2(p)) ? 1 : 0;
}
// pairing check: Π e(P_i, Q_i) == 1 ; g1: n×64 B, g2: n×128 B
int emu_pairing_check(const uint8_t* g1, const uint8_t* g2, int n) {
    static PairingTables pt; static bool init = false;
    if (!init) { pairing_tables_init(pt); init = true; }
    Fq12 f = Fq12::one();
    for (int i = 0; i < n; i++) f = f * miller_loop(&pt, ld_g2(g2 + 128 * i), ld_g1(g1 + 64 * i));
    return final_exponentiation(&pt, f) == Fq12::one() ? 1 : 0;
}
// 1 if the addition-chain final exponentiation and the generic one agree on "is the result one" for Π e(P_i,Q_i),
// and additionally the chain result is multiplicative: FE(f·g) == FE(f)·FE(g)
int emu_final_exp_consistency(const uint8_t* g1, const uint8_t* g2, int n) {
    static PairingTables pt; static bool init = false;
    if (!init) { pairing_tables_init(pt); init = true; }
    Fq12 f = Fq12::one(), g = Fq12::one();
    for (int i = 0; i < n; i++) {
        Fq12 m = miller_loop(&pt, ld_g2(g2 + 128 * i), ld_g1(g1 + 64 * i));
        f = f * m;
        if (i == 0) g = m;
    }
    bool one_chain = final_exponentiation(&pt, f) == Fq12::one();
    bool one_generic = final_exponentiation_generic(&pt, f) == Fq12::one();
    Fq12 a = final_exponentiation(&pt, f * g), b = final_exponentiation(&pt, f) * final_exponentiation(&pt, g);
    // fixed-point Miller loop (precomputed lines) must reproduce the generic one
    static FixedLines fl;
    precompute_lines(&pt, ld_g2(g2), fl);
    bool fixed_ok = miller_loop_fixed(fl.lam, fl.c, ld_g1(g1)) == g;
    if (!fixed_ok) return 0;
    // frobenius1 applied twice must equal frobenius2
    bool frob_ok = frobenius1(&pt, frobenius1(&pt, g)) == frobenius2(&pt, g);
    return (one_chain == one_generic ? 1 : 0) | (a == b ? 2 : 0) | (frob_ok ? 4 : 0) | (one_chain ? 8 : 0);
}
// the verifier's fast paths against the plain ones (bit flags; all must be set):
//  1  projective Miller loop == affine Miller loop after the final exponentiation
//  2  mul_by_034 == dense product with the same line            4  mul_by_line_fq == dense product
//  8  cyclotomic_sqr == sqr on an element of the cyclotomic subgroup
// 16  the merged Groth16 loop == product of three separate loops after the final exponentiation
int emu_pairing_fast_paths(const uint8_t* g1, const uint8_t* g2) {   // g1: 3 × 64 B (A, V, C), g2: 3 × 128 B (B, G, D)
    static PairingTables pt; static bool init = false;
    if (!init) { pairing_tables_init(pt); init = true; }
    int flags = 0;
    const G1Affine A = ld_g1(g1), V = ld_g1(g1 + 64), C = ld_g1(g1 + 128);
    const G2Affine B = ld_g2(g2), G = ld_g2(g2 + 128), D = ld_g2(g2 + 256);
    const Fq12 m_aff = miller_loop(&pt, B, A), m_proj = miller_loop_proj(&pt, B, A);
    if (final_exponentiation(&pt, m_aff) == final_exponentiation(&pt, m_proj)) flags |= 1;
    {   // sparse products on a dense, non-trivial element
        const Fq12 f = m_aff;
        const Fq2 l0 = B.x * B.y, l3 = B.y.sqr(), l4 = B.x.sqr();
        Fq12 dense;
        dense.c0 = {l0, Fq2::zero(), Fq2::zero()};
        dense.c1 = {l3, l4, Fq2::zero()};
        if (f.mul_by_034(l0, l3, l4) == f * dense) flags |= 2;
        dense.c0 = {Fq2{A.y, Fq::zero()}, Fq2::zero(), Fq2::zero()};
        if (f.mul_by_line_fq(A.y, l3, l4) == f * dense) flags |= 4;
    }
    {
        Fq12 t = m_aff.conj() * m_aff.inv();
        t = frobenius2(&pt, t) * t;   // easy part done: t is in the cyclotomic subgroup
        if (t.cyclotomic_sqr() == t.sqr() && t.cyclotomic_sqr().cyclotomic_sqr() == t.sqr().sqr()) flags |= 8;
    }
    {
        static FixedLines lg, ld_;
        precompute_lines(&pt, G, lg);
        precompute_lines(&pt, D, ld_);
        const Fq12 merged = miller_loop_groth16(&pt, B, A, lg.lam, lg.c, V, ld_.lam, ld_.c, C);
        const Fq12 separate = miller_loop(&pt, B, A) * miller_loop(&pt, G, V) * miller_loop(&pt, D, C);
        if (final_exponentiation(&pt, merged) == final_exponentiation(&pt, separate)) flags |= 16;
    }
    return flags;
}
// both G2 membership tests on one twist point: bit 0 = ψ(P) = [6x²]P, bit 1 = the one-multiple test
int emu_g2_subgroup_both(const uint8_t* p) {
    static PairingTables pt; static bool init = false;
    if (!init) { pairing_tables_init(pt); init = true; }
    const G2Affine P = ld_g2(p);
    return (g2_in_subgroup_6x2(&pt, P) ? 1 : 0) | (g2_in_subgroup(&pt, P) ? 2 : 0);
}
// witness-graph VM: one op on canonical values
int emu_vm_duo(int op, const uint8_t* a, const uint8_t* b, uint8_t* out) {
    Fr r;
    bool ok = vm_eval_duo(op, ld<Fr>(a), ld<Fr>(b), r);
    st(out, r);
    return ok ? 1 : 0;
}
}

// tree configuration parser (host_util.hpp parse_tree_config ↔ rln/src/pm_tree_adapter.rs:139-174)
extern "C" int emu_parse_tree_config(const char* json, char* path_out, size_t path_cap, uint64_t* nums /* temporary, has_path, cache, flush_ms, low_space,
                                     compression, has_depth, depth */, char* err_out, size_t err_cap) {
    try {
        TreeConfig c = parse_tree_config(json);
        snprintf(path_out, path_cap, "%s", c.path.c_str());
        nums[0] = c.temporary; nums[1] = c.has_path; nums[2] = c.cache_capacity; nums[3] = c.flush_every_ms; nums[4] = c.low_space;
        nums[5] = c.use_compression; nums[6] = c.has_depth; nums[7] = c.tree_depth;
        return 0;
    } catch (const std::exception& e) {
        snprintf(err_out, err_cap, "%s", e.what());
        return 1;
    }
}
