// Lane-parallel Groth16 verification: the "pairing VM".
//
// rln/src/protocol/proof.rs:856-894 (verify_zk_proof → ark-groth16 verify_proof) is, for ONE proof, ≈ 36 000 dependent-looking Fq
// products; one GPU thread needs 22 ms for them, which is slower than a CPU.  Almost all of that work is data-INDEPENDENT
// straight-line arithmetic (square-root chains of the decompression, the G2 membership test, the Miller loop, the final
// exponentiation), and a warp executes 32 lanes for the price of one.  So the host traces the whole verification ONCE per
// verifying key into a dataflow program over Fq whose only operation is
//
//        slot[dst] = Σ_{t<N} ± 2^{sh_t} · slot[a_t] · slot[b_t]          (one Montgomery reduction per sum, N ≤ 8)
//
// schedules it onto NW warps × 32 lanes (list scheduling by critical path, one "level" = one such operation on every lane, a
// barrier between levels) and assigns shared-memory slots; k_verify_vm (k_verify_vm.cu) interprets the levels for one proof per
// CTA.  Towers are flattened so that every level is as wide as possible: Fq12 = Fq2[w]/(w⁶ − ξ) with schoolbook products
// (12 outputs × 12 terms instead of Karatsuba's dependent additions), the ξ-multiples an operand needs for the wrap-around are
// produced as extra outputs of the level that produced the operand, sums of more than NMAX terms are split over the lane pair
// (l, l^16) and recombined with shuffles.  The few data-dependent decisions (root selection and sign bits of the
// decompression, vk_x from the public inputs, the final comparisons) are "special" levels run by hand-written code
// (verify_vm_special.cuh); anything unusual (points at infinity, exceptional additions) is reported as FALLBACK and re-run by
// the one-thread-per-proof kernel, which has complete formulas.
//
// This header is host code (the tracer / scheduler) plus the record format shared with the kernel and with the host
// interpreter used by tests/host_emul.
#pragma once
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <utility>
#include <vector>

#include "pairing_constants.hpp"
#include "tower.cuh"

namespace zk {
namespace pvm {

constexpr int NW = 4;                 // VM warps per proof
constexpr int LANES = 32 * NW;
constexpr int NMAX = 8;               // terms per lane and level
constexpr int REC_WORDS = 2 + NMAX;   // per lane and level: w0, w1, NMAX term words; stored [level][word][lane]
// w0: dst slot (12 bits) | STORE << 12 | COMBINE << 13 (add the result of lane ^ 16) | COMBINE4 << 14 (then that of lane ^ 8):
//     a sum split over two lanes sits on (l, l+16), one split over four on (l, l+8, l+16, l+24), l < 8
// w1: N (4 bits: terms of this warp in this level) | nsub << 4 (2 bits: conditional subtractions) | any-combine << 6 | any-combine4 << 7 | special id << 8
//     (8 bits) | exponent index << 16 (SP_EXP)
// term: slot a (12 bits) | slot b << 12 | negate << 24 | shift << 25 (2 bits: the a operand is scaled by 2^shift); SP_EXP: term 0
//     = source slot | first scratch slot << 12
// a special level carries its argument slots in w0 of lanes 0..n_args−1 of warp 0
constexpr u32 W0_STORE = 1u << 12, W0_COMBINE = 1u << 13, W0_COMBINE4 = 1u << 14;
enum SpecialId : u32 { SP_NONE = 0, SP_VKX = 1, SP_SELECT = 2, SP_FINAL = 3, SP_EXP = 4 };
// an SP_EXP level raises up to EXP_LANES values to one of the three public exponents of the verification, each lane on its own
// (a chain of ≈ 320 dependent products would otherwise be ≈ 320 levels of one product each, and a level costs ≈ 1 000 cycles
// before its first product): lane i < EXP_LANES of warp 0 with W0_STORE set computes slot[dst] = slot[src]^e, src in term word 0,
// the exponent's index in bits 16.. of w1; its window table lives in the scratch block i at the top of the slot file
constexpr int EXP_LANES = 4, EXP_TABLE = 16;
enum ExpId : u32 { EXP_SQRT = 0, EXP_INV_SQRT = 1, EXP_INV = 2 };   // (q+1)/4, (q−3)/4, q−2
enum Status : u32 { ST_RUNNING = 0, ST_VALID = 1, ST_INVALID = 4, ST_MALFORMED = 2, ST_FALLBACK = 3 };   // ok[] codes: 1 valid, 0 invalid, 2 malformed, 3 re-run
// pinned slots
enum Pinned : int {
    S_ZERO = 0, S_ONE, S_RAW1, S_R2, S_XA, S_XC, S_XB0, S_XB1, S_VX, S_VY, S_VZZ, S_VZZZ, S_FIRST_CONST
};
constexpr int MAX_SLOTS = 4096;

struct Program {
    std::vector<u32> code;        // n_levels × REC_WORDS × LANES
    std::vector<Fq> consts;       // image of slots [0, n_const): pinned constants (inputs are zero here)
    u32 n_levels = 0, n_const = 0, n_slots = 0;   // n_slots includes the EXP_LANES × EXP_TABLE scratch slots at the top
    u32 exps[3][8] = {};          // the public exponents, little-endian words
    double est_cycles = 0;        // the scheduler's cost model, for reports
    u32 n_nodes = 0;
};

// ================================================================================================== symbolic values
// a value = integer linear combination of materialised nodes
struct Val {
    std::vector<std::pair<int, int>> t;   // (node, coefficient), sorted by node, no zero coefficients
    bool empty() const { return t.empty(); }
};
inline Val operator+(const Val& x, const Val& y) {
    Val r;
    size_t i = 0, j = 0;
    while (i < x.t.size() || j < y.t.size()) {
        if (j == y.t.size() || (i < x.t.size() && x.t[i].first < y.t[j].first)) r.t.push_back(x.t[i++]);
        else if (i == x.t.size() || y.t[j].first < x.t[i].first) r.t.push_back(y.t[j++]);
        else { const int c = x.t[i].second + y.t[j].second; if (c) r.t.push_back({x.t[i].first, c}); i++; j++; }
    }
    return r;
}
inline Val operator*(const Val& x, int k) {
    Val r;
    if (k) for (auto& e : x.t) r.t.push_back({e.first, e.second * k});
    return r;
}
inline Val operator-(const Val& x) { return x * -1; }
inline Val operator-(const Val& x, const Val& y) { return x + (y * -1); }

struct S2 {   // Fq2
    Val a, b;
    bool empty() const { return a.empty() && b.empty(); }
};
inline S2 operator+(const S2& x, const S2& y) { return {x.a + y.a, x.b + y.b}; }
inline S2 operator-(const S2& x, const S2& y) { return {x.a - y.a, x.b - y.b}; }
inline S2 operator*(const S2& x, int k) { return {x.a * k, x.b * k}; }
inline S2 conj(const S2& x) { return {x.a, x.b * -1}; }
inline S2 mul_xi(const S2& x) { return {x.a * 9 - x.b, x.b * 9 + x.a}; }   // lazily: coefficients 9 (materialise before multiplying)

struct Term { int a, b, coef; };

struct Node {
    enum Kind : uint8_t { PINNED, DOT, SPECIAL, SPECIAL_OUT, EXP } kind = DOT;
    std::vector<Term> terms;      // DOT (after coefficient splitting: |coef| ∈ {1, 2, 4})
    std::vector<int> deps;        // SPECIAL: argument nodes (inputs then outputs); SPECIAL_OUT: the special
    int special_id = 0;           // SPECIAL: which; EXP: the exponent's index
    int slot = -1;                // PINNED / pinned SPECIAL_OUT: fixed
    int split = 1;                // terms spread over 1, 2 or 4 lanes
    int n = 0, nsub = 1;          // terms per lane, conditional subtractions
    int level = -1, lane = -1;
    double prio = 0;
    int last_use = -1;
};

struct Builder {
    std::vector<Node> nodes;
    std::vector<Fq> consts;       // slot image
    std::map<int, int> kconst_;   // small integer → node
    int split_threshold = NMAX;   // sums with more terms than this go to a lane pair
    int one_, zero_;

    Builder() {
        consts.assign(S_FIRST_CONST, Fq::zero());
        for (int s = 0; s < S_FIRST_CONST; s++) { Node n; n.kind = Node::PINNED; n.slot = s; n.level = -1; nodes.push_back(n); }
        consts[S_ONE] = Fq::one();
        consts[S_RAW1] = Fq::zero(); consts[S_RAW1].l[0] = 1;
        consts[S_R2] = Fq::rsquared();
        // S_VX… are written by the vk_x warp: as nodes they are re-declared as outputs of SP_VKX by the program
        one_ = S_ONE; zero_ = S_ZERO;
    }
    Val v(int node) const { Val r; r.t.push_back({node, 1}); return r; }
    Val one() const { return v(one_); }
    int pin(const Fq& value) {
        Node n; n.kind = Node::PINNED; n.slot = (int)consts.size();
        consts.push_back(value);
        nodes.push_back(n);
        return (int)nodes.size() - 1;
    }
    Val cfq(const Fq& value) { return v(pin(value)); }
    S2 cfq2(const Fq2& value) { return {cfq(value.a), cfq(value.b)}; }
    Val kconst(int k) {   // Montgomery form of a small positive integer
        auto it = kconst_.find(k);
        if (it != kconst_.end()) return v(it->second);
        Fq x = Fq::zero(), o = Fq::one();
        for (int i = 0; i < k; i++) x = x + o;
        const int id = pin(x);
        kconst_[k] = id;
        return v(id);
    }
    // Σ xᵢ·yᵢ, expanded over the linear combinations, as ONE node
    Val dot(const std::vector<std::pair<Val, Val>>& prods) {
        std::map<std::pair<int, int>, int> acc;
        for (auto& pr : prods)
            for (auto& x : pr.first.t)
                for (auto& y : pr.second.t) {
                    std::pair<int, int> key = std::minmax(x.first, y.first);
                    acc[key] += x.second * y.second;
                }
        Node n;
        n.kind = Node::DOT;
        int W = 0;
        for (auto& e : acc) {
            int c = e.second;
            if (!c) continue;
            const int sgn = c < 0 ? -1 : 1;
            c = c < 0 ? -c : c;
            W += c;
            for (int p2 = 4; p2 >= 1; p2 >>= 1)
                while (c >= p2) { n.terms.push_back({e.first.first, e.first.second, sgn * p2}); c -= p2; }
        }
        if (n.terms.empty()) return Val{};
        const int nt = (int)n.terms.size();
        if (nt > 4 * NMAX) throw std::runtime_error("pvm: sum too long: " + std::to_string(nt));
        n.split = nt <= split_threshold ? 1 : nt <= 2 * split_threshold ? 2 : 4;
        if ((nt + n.split - 1) / n.split > NMAX) n.split *= 2;
        n.n = (nt + n.split - 1) / n.split;
        if (n.split > 1) {   // balance the weight of the parts: heavy terms are dealt round robin
            std::stable_sort(n.terms.begin(), n.terms.end(), [](const Term& a, const Term& b) { return std::abs(a.coef) > std::abs(b.coef); });
            std::vector<std::vector<Term>> part(n.split);
            for (size_t i = 0; i < n.terms.size(); i++) part[i % n.split].push_back(n.terms[i]);
            n.terms.clear();
            W = 0;
            for (auto& h : part) {
                h.resize(n.n, Term{zero_, zero_, 1});
                int w = 0;
                for (auto& t : h) if (t.a != zero_) w += std::abs(t.coef);
                W = std::max(W, w);
                n.terms.insert(n.terms.end(), h.begin(), h.end());
            }
        }
        // (W·p² + R·p)/R = (0.18903·W + 1)·p must come below p after nsub conditional subtractions
        n.nsub = W <= 5 ? 1 : W <= 10 ? 2 : W <= 15 ? 3 : 0;
        if (!n.nsub) throw std::runtime_error("pvm: weight too large: " + std::to_string(W));
        nodes.push_back(n);
        return v((int)nodes.size() - 1);
    }
    Val mul(const Val& x, const Val& y) { return dot({{x, y}}); }
    // a linear combination as a node of its own (coefficients beyond 4 go through a constant slot)
    Val mat(const Val& x) {
        if (x.t.size() == 1 && x.t[0].second == 1) return x;
        std::vector<std::pair<Val, Val>> pr;
        for (auto& e : x.t) {
            const int c = e.second < 0 ? -e.second : e.second;
            if (c <= 4 && c != 3) pr.push_back({v(e.first) * e.second, one()});
            else pr.push_back({v(e.first) * (e.second < 0 ? -1 : 1), kconst(c)});
        }
        return dot(pr);
    }
    S2 mat(const S2& x) { return {mat(x.a), mat(x.b)}; }
    // a special level: arguments = inputs (materialised) followed by n_out fresh outputs (or pinned ones)
    std::vector<Val> special(int id, std::vector<Val> inputs, int n_out, const std::vector<int>& pinned_out = {}) {
        Node s;
        s.kind = Node::SPECIAL;
        s.special_id = id;
        for (auto& in : inputs) {
            Val m = in.empty() ? v(zero_) : mat(in);
            s.deps.push_back(m.t[0].first);
        }
        const int sid = (int)nodes.size();
        nodes.push_back(s);
        std::vector<Val> outs;
        for (int i = 0; i < n_out; i++) {
            Node o;
            o.kind = Node::SPECIAL_OUT;
            o.deps.push_back(sid);
            if (i < (int)pinned_out.size()) o.slot = pinned_out[i];
            nodes.push_back(o);
            nodes[sid].deps.push_back((int)nodes.size() - 1);
            outs.push_back(v((int)nodes.size() - 1));
        }
        nodes[sid].n = (int)inputs.size();   // number of inputs among deps
        return outs;
    }

    // ---------------------------------------------------------------------------------------------- Fq2 helpers
    struct Acc2 {   // Σ of Fq2 products and linear parts, emitted as two nodes
        Builder& B;
        std::vector<std::pair<Val, Val>> re, im;
        explicit Acc2(Builder& b) : B(b) {}
        void add(const S2& x, const S2& y, int k = 1) {
            if (x.empty() || y.empty()) return;
            re.push_back({x.a * k, y.a}); re.push_back({x.b * -k, y.b});
            im.push_back({x.a * k, y.b}); im.push_back({x.b * k, y.a});
        }
        void add_fq(const S2& x, const Val& s, int k = 1) { re.push_back({x.a * k, s}); im.push_back({x.b * k, s}); }
        void lin(const S2& x, int k = 1) { add_fq(x, B.one(), k); }
        S2 emit() { return {B.dot(re), B.dot(im)}; }
    };
    S2 mul(const S2& x, const S2& y, int k = 1) { Acc2 a(*this); a.add(x, y, k); return a.emit(); }
    S2 mul_fq(const S2& x, const Val& s, int k = 1) { Acc2 a(*this); a.add_fq(x, s, k); return a.emit(); }
    S2 one2() const { return {one(), Val{}}; }

    // ---------------------------------------------------------------------------------------------- fixed exponents
    // a^e for one of the public exponents: one node, evaluated by a lane on its own in an SP_EXP level
    Val pow_fixed(const Val& a_in, int exp_id) {
        const Val a = mat(a_in);
        Node n;
        n.kind = Node::EXP;
        n.special_id = exp_id;
        n.deps.push_back(a.t[0].first);
        n.n = 1;
        nodes.push_back(n);
        return v((int)nodes.size() - 1);
    }

    // ---------------------------------------------------------------------------------------------- Fq12 = Fq2[w]/(w⁶ − ξ)
    struct S12 {
        S2 c[6];     // coefficient of w^k  (tower.cuh: k = 0,2,4 ↔ c0.c0,c0.c1,c0.c2 and k = 1,3,5 ↔ c1.c0,c1.c1,c1.c2)
        S2 x[6];     // ξ·c[k], materialised (k = 1…5), when has_x
        S2 x2[6];    // ξ²·c[k]
        bool has_x = false, has_x2 = false;
    };
    S12 one12() { S12 r; r.c[0] = one2(); return r; }
    void ensure_x(S12& a) {
        if (a.has_x) return;
        for (int k = 0; k < 6; k++) if (!a.c[k].empty()) a.x[k] = mat(mul_xi(a.c[k]));
        a.has_x = true;
    }
    void ensure_x2(S12& a) {
        ensure_x(a);
        if (a.has_x2) return;
        for (int k = 1; k < 6; k++) if (!a.x[k].empty()) a.x2[k] = mat(mul_xi(a.x[k]));
        a.has_x2 = true;
    }
    // a·b; b needs its ξ-multiples (and the ξ² ones if the result is to carry its own ξ-multiples)
    S12 mul12(const S12& a, const S12& b, bool want_x) {
        if (!b.has_x || (want_x && !b.has_x2)) throw std::runtime_error("pvm: mul12 operand without its xi multiples");
        S12 r;
        for (int k = 0; k < 6; k++) {
            Acc2 acc(*this);
            for (int i = 0; i < 6; i++) { const int m = k - i; acc.add(a.c[i], m >= 0 ? b.c[m] : b.x[m + 6]); }
            r.c[k] = acc.emit();
        }
        if (want_x) {
            for (int k = 0; k < 6; k++) {
                Acc2 acc(*this);
                for (int i = 0; i < 6; i++) { const int m = k - i; acc.add(a.c[i], m >= 0 ? b.x[m] : b.x2[m + 6]); }
                r.x[k] = acc.emit();
            }
            r.has_x = true;
        }
        return r;
    }
    // a² from (a, ξa): symmetric sums, coefficient 2 on the mixed products
    S12 sqr12(const S12& a) {
        if (!a.has_x) throw std::runtime_error("pvm: sqr12 operand without its xi multiples");
        S12 r;
        for (int k = 0; k < 6; k++) {
            Acc2 acc(*this);
            for (int i = 0; i < 6; i++)
                for (int j = i; j < 6; j++) {
                    if (i + j == k) acc.add(a.c[i], a.c[j], i == j ? 1 : 2);
                    else if (i + j == k + 6) acc.add(a.c[i], a.x[j], i == j ? 1 : 2);
                }
            r.c[k] = acc.emit();
        }
        return r;
    }
    S12 conj12(const S12& a) {   // a^(q⁶): w ↦ −w
        S12 r = a;
        for (int k = 1; k < 6; k += 2) { r.c[k] = a.c[k] * -1; r.x[k] = a.x[k] * -1; r.x2[k] = a.x2[k] * -1; }
        return r;
    }
    // Granger–Scott squaring in the cyclotomic subgroup (tower.cuh Fq12::cyclotomic_sqr), one level: the ξ-multiples of the
    // result are bilinear in (a, ξa) as well.  z0..z5 ↔ w⁰, w³, w¹, w⁴, w², w⁵.
    S12 cyc_sqr(const S12& a) {
        if (!a.has_x) throw std::runtime_error("pvm: cyc_sqr operand without its xi multiples");
        static const int ZI[6] = {0, 3, 1, 4, 2, 5};
        auto z = [&](int i) -> const S2& { return a.c[ZI[i]]; };
        auto xz = [&](int i) -> const S2& { return a.x[ZI[i]]; };
        S12 r;
        auto out = [&](int i, bool x) -> S2& { return x ? r.x[ZI[i]] : r.c[ZI[i]]; };
        { Acc2 c(*this); c.add(z(0), z(0), 3); c.add(z(1), xz(1), 3); c.lin(z(0), -2); out(0, false) = c.emit(); }
        { Acc2 c(*this); c.add(z(0), z(1), 6); c.lin(z(1), 2); out(1, false) = c.emit(); }
        { Acc2 c(*this); c.add(z(4), xz(5), 6); c.lin(z(2), 2); out(2, false) = c.emit(); }
        { Acc2 c(*this); c.add(z(4), z(4), 3); c.add(z(5), xz(5), 3); c.lin(z(3), -2); out(3, false) = c.emit(); }
        { Acc2 c(*this); c.add(z(2), z(2), 3); c.add(z(3), xz(3), 3); c.lin(z(4), -2); out(4, false) = c.emit(); }
        { Acc2 c(*this); c.add(z(2), z(3), 6); c.lin(z(5), 2); out(5, false) = c.emit(); }
        { Acc2 c(*this); c.add(z(0), xz(0), 3); c.add(xz(1), xz(1), 3); c.lin(xz(0), -2); out(0, true) = c.emit(); }
        { Acc2 c(*this); c.add(z(0), xz(1), 6); c.lin(xz(1), 2); out(1, true) = c.emit(); }
        { Acc2 c(*this); c.add(xz(4), xz(5), 6); c.lin(xz(2), 2); out(2, true) = c.emit(); }
        { Acc2 c(*this); c.add(z(4), xz(4), 3); c.add(xz(5), xz(5), 3); c.lin(xz(3), -2); out(3, true) = c.emit(); }
        { Acc2 c(*this); c.add(z(2), xz(2), 3); c.add(xz(3), xz(3), 3); c.lin(xz(4), -2); out(4, true) = c.emit(); }
        { Acc2 c(*this); c.add(z(2), xz(3), 6); c.lin(xz(5), 2); out(5, true) = c.emit(); }
        r.has_x = true;
        return r;
    }
    // Frobenius maps (tower.cuh frobenius1 / frobenius2); constants are pinned once
    S2 F1_[6], XF1_[6]; Val F2_[6]; bool frob_init_ = false;
    void init_frob(const PairingTables& pt) {
        const Fq2 xi = {Fq::from_u32(9), Fq::from_u32(1)};
        for (int k = 0; k < 6; k++) { F1_[k] = cfq2(pt.frob1[k]); XF1_[k] = cfq2(pt.frob1[k] * xi); F2_[k] = cfq(pt.frob2[k]); }
        frob_init_ = true;
    }
    S12 frob1(const S12& a, bool want_x) {
        S12 r;
        for (int k = 0; k < 6; k++) r.c[k] = mul(conj(a.c[k]), F1_[k]);
        if (want_x) { for (int k = 0; k < 6; k++) r.x[k] = mul(conj(a.c[k]), XF1_[k]); r.has_x = true; }
        return r;
    }
    S12 frob2(const S12& a, bool want_x) {
        S12 r;
        for (int k = 0; k < 6; k++) r.c[k] = k == 0 ? a.c[k] : mul_fq(a.c[k], F2_[k]);
        if (want_x) {
            if (!a.has_x) throw std::runtime_error("pvm: frob2 wants xi multiples of its operand");
            for (int k = 0; k < 6; k++) r.x[k] = k == 0 ? a.x[k] : mul_fq(a.x[k], F2_[k]);
            r.has_x = true;
        }
        return r;
    }

    // ---------------------------------------------------------------------------------------------- scheduling
    static constexpr double EXP_COST = 250000.0;   // ≈ 254 squarings + 78 products of one lane
    // measured on a B200 (scratch/vm_trace.py, profiles/r02m_*): ≈ 1 350 cycles per level + 570 per term
    static double level_cost(int n, int nsub, bool comb) { return 1350.0 + 570.0 * n + 70.0 * (nsub - 1) + (comb ? 250.0 : 0.0); }
    Program schedule(int window = 6000) {
        const int NN = (int)nodes.size();
        // consumers / priorities (nodes are in topological order)
        std::vector<std::vector<int>> deps(NN);
        for (int i = 0; i < NN; i++) {
            Node& n = nodes[i];
            if (n.kind == Node::DOT) {
                for (auto& t : n.terms) { deps[i].push_back(t.a); deps[i].push_back(t.b); }
            } else if (n.kind == Node::SPECIAL) {
                for (int k = 0; k < n.n; k++) deps[i].push_back(n.deps[k]);
            } else if (n.kind == Node::SPECIAL_OUT || n.kind == Node::EXP) {
                deps[i].push_back(n.deps[0]);
            }
            std::sort(deps[i].begin(), deps[i].end());
            deps[i].erase(std::unique(deps[i].begin(), deps[i].end()), deps[i].end());
        }
        // dead values (multiples nobody asked for, the tail of running state) are not scheduled
        std::vector<char> live(NN, 0);
        for (int i = NN - 1; i >= 0; i--) {
            if (nodes[i].kind == Node::SPECIAL || nodes[i].kind == Node::PINNED) live[i] = 1;
            if (nodes[i].kind == Node::SPECIAL_OUT) live[i] = 1;   // written by hand-written code whether read or not
            if (live[i]) for (int d : deps[i]) live[d] = 1;
        }
        for (int i = NN - 1; i >= 0; i--) {
            Node& n = nodes[i];
            if (!live[i]) continue;
            const double own = n.kind == Node::DOT ? level_cost(n.n, n.nsub, n.split > 1) : n.kind == Node::SPECIAL ? 2000.0 : n.kind == Node::EXP ? EXP_COST : 0.0;
            n.prio += own;
            for (int d : deps[i]) nodes[d].prio = std::max(nodes[d].prio, n.prio);
        }
        std::vector<int> remaining(NN, 0);
        std::vector<std::vector<int>> users(NN);
        for (int i = 0; i < NN; i++) if (live[i]) for (int d : deps[i]) { users[d].push_back(i); }
        std::vector<int> avail_level(NN, 0);   // first level at which the node may run
        std::vector<int> ready;
        for (int i = 0; i < NN; i++) {
            if (!live[i]) { nodes[i].level = -2; continue; }
            for (int d : deps[i]) if (nodes[d].kind != Node::PINNED) remaining[i]++;
            if (nodes[i].kind == Node::PINNED) { nodes[i].level = -1; continue; }
            if (!remaining[i]) ready.push_back(i);
        }
        struct Level { int special = -1; int exp_id = -1; int lane_node[LANES]; int lane_half[LANES]; };
        std::vector<Level> levels;
        int n_done = 0, n_todo = 0, lowest_open = 0;
        for (int i = 0; i < NN; i++) if (nodes[i].kind != Node::PINNED && live[i]) n_todo++;
        auto finish = [&](int i, int lvl) {
            nodes[i].level = lvl;
            n_done++;
            for (int u : users[i]) {
                avail_level[u] = std::max(avail_level[u], lvl + 1);
                if (--remaining[u] == 0) ready.push_back(u);
            }
        };
        double est = 0;
        while (n_done < n_todo) {
            const int lvl = (int)levels.size();
            while (lowest_open < NN && (nodes[lowest_open].kind == Node::PINNED || nodes[lowest_open].level >= 0 || !live[lowest_open])) lowest_open++;
            // special outputs become available right after their special
            {
                bool again = true;
                while (again) {
                    again = false;
                    for (size_t r = 0; r < ready.size(); r++) {
                        const int i = ready[r];
                        if (nodes[i].kind == Node::SPECIAL_OUT) {
                            ready.erase(ready.begin() + r);
                            finish(i, nodes[nodes[i].deps[0]].level);
                            again = true;
                            break;
                        }
                    }
                }
            }
            std::vector<int> cand;
            for (int i : ready) if (avail_level[i] <= lvl && i < lowest_open + window) cand.push_back(i);
            if (cand.empty()) {
                if (ready.empty()) throw std::runtime_error("pvm: scheduler stuck");
                // nothing may run yet in the window: widen it for this level
                for (int i : ready) if (avail_level[i] <= lvl) cand.push_back(i);
                if (cand.empty()) throw std::runtime_error("pvm: scheduler stuck (levels)");
            }
            std::sort(cand.begin(), cand.end(), [&](int a, int b) { return nodes[a].prio != nodes[b].prio ? nodes[a].prio > nodes[b].prio : a < b; });
            Level L;
            for (int l = 0; l < LANES; l++) { L.lane_node[l] = -1; L.lane_half[l] = 0; }
            std::vector<int> placed;
            if (nodes[cand[0]].kind == Node::SPECIAL) {
                L.special = cand[0];
                placed.push_back(cand[0]);
                est += 2000.0;
            } else if (nodes[cand[0]].kind == Node::EXP) {   // every ready power with the same exponent shares the level
                L.exp_id = nodes[cand[0]].special_id;
                for (int i : cand)
                    if (nodes[i].kind == Node::EXP && nodes[i].special_id == L.exp_id && (int)placed.size() < EXP_LANES) {
                        L.lane_node[placed.size()] = i;
                        nodes[i].lane = (int)placed.size();
                        placed.push_back(i);
                    }
                est += EXP_COST;
            } else {
                int warp_n[NW], warp_cap[NW];
                for (int w = 0; w < NW; w++) { warp_n[w] = 0; warp_cap[w] = 0; }
                const int top_n = nodes[cand[0]].n;
                for (int i : cand) {
                    Node& n = nodes[i];
                    if (n.kind != Node::DOT) continue;
                    if (n.n > std::max(top_n, 2)) continue;   // a wider sum would slow the level of the critical node down
                    int best_w = -1, best_l = -1;
                    for (int w = 0; w < NW && best_w < 0; w++) {
                        if (warp_cap[w] && n.n > warp_cap[w]) continue;
                        for (int l = 0; l < 32; l++) {
                            const int gl = w * 32 + l;
                            if (L.lane_node[gl] >= 0) continue;
                            if (n.split == 2) { if (l >= 16 || L.lane_node[gl + 16] >= 0) continue; }
                            if (n.split == 4) { if (l >= 8 || L.lane_node[gl + 8] >= 0 || L.lane_node[gl + 16] >= 0 || L.lane_node[gl + 24] >= 0) continue; }
                            best_w = w; best_l = l;
                            break;
                        }
                    }
                    if (best_w < 0) continue;
                    const int gl = best_w * 32 + best_l;
                    L.lane_node[gl] = i; L.lane_half[gl] = 0;
                    if (n.split == 2) { L.lane_node[gl + 16] = i; L.lane_half[gl + 16] = 1; }
                    if (n.split == 4) for (int h = 1; h < 4; h++) { L.lane_node[gl + 8 * h] = i; L.lane_half[gl + 8 * h] = h; }
                    if (!warp_cap[best_w]) warp_cap[best_w] = std::max(n.n, 2);
                    warp_n[best_w] = std::max(warp_n[best_w], n.n);
                    n.lane = gl;
                    placed.push_back(i);
                }
                double c = 0;
                for (int w = 0; w < NW; w++) {
                    if (!warp_n[w]) continue;
                    int nsub = 1; bool comb = false;
                    for (int l = 0; l < 32; l++) if (L.lane_node[w * 32 + l] >= 0) { nsub = std::max(nsub, nodes[L.lane_node[w * 32 + l]].nsub); comb |= nodes[L.lane_node[w * 32 + l]].split > 1; }
                    c = std::max(c, level_cost(warp_n[w], nsub, comb));
                }
                est += c;
            }
            if (placed.empty()) throw std::runtime_error("pvm: empty level");
            if (getenv("PVM_TRACE")) {
                int used = 0, mx = 0;
                for (int l = 0; l < LANES; l++) if (L.lane_node[l] >= 0) { used++; mx = std::max(mx, nodes[L.lane_node[l]].n); }
                fprintf(stderr, "L%d special=%d lanes=%d maxN=%d first=%d est=%.0f\n", lvl, L.special >= 0 ? nodes[L.special].special_id : 0, used, mx, placed[0], est);
            }
            for (int i : placed) ready.erase(std::find(ready.begin(), ready.end(), i));
            levels.push_back(L);
            for (int i : placed) finish(i, lvl);
        }
        // ---- slots: a node's slot is free again one level after its last reader
        const int n_levels = (int)levels.size();
        for (int i = 0; i < NN; i++) {
            Node& n = nodes[i];
            if (n.kind == Node::PINNED || !live[i]) continue;
            for (int d : deps[i]) nodes[d].last_use = std::max(nodes[d].last_use, n.level);
            if (n.kind == Node::SPECIAL) for (size_t k = n.n; k < n.deps.size(); k++) nodes[n.deps[k]].last_use = std::max(nodes[n.deps[k]].last_use, n.level);
        }
        const int first_dyn = (int)consts.size();
        std::vector<int> free_slots;
        int next_slot = first_dyn;
        std::vector<std::vector<int>> born(n_levels), dies(n_levels + 1);
        for (int i = 0; i < NN; i++) {
            Node& n = nodes[i];
            if (n.kind == Node::PINNED || n.kind == Node::SPECIAL || n.slot >= 0 || !live[i]) continue;
            born[n.level].push_back(i);
            dies[std::min(n_levels, std::max(n.last_use, n.level) + 1)].push_back(i);
        }
        for (int l = 0; l < n_levels; l++) {
            for (int i : dies[l]) free_slots.push_back(nodes[i].slot);
            for (int i : born[l]) {
                if (!free_slots.empty()) { nodes[i].slot = free_slots.back(); free_slots.pop_back(); }
                else nodes[i].slot = next_slot++;
            }
        }
        const int scratch_base = next_slot;
        next_slot += EXP_LANES * EXP_TABLE;
        if (next_slot > MAX_SLOTS) throw std::runtime_error("pvm: out of slots: " + std::to_string(next_slot));
        // ---- records
        Program P;
        P.n_levels = (u32)n_levels;
        P.n_const = (u32)first_dyn;
        P.n_slots = (u32)next_slot;
        P.consts = consts;
        P.est_cycles = est;
        P.n_nodes = (u32)NN;
        P.code.assign((size_t)n_levels * REC_WORDS * LANES, 0);
        const u32 zero_term = (u32)S_ZERO | ((u32)S_ZERO << 12);
        for (int l = 0; l < n_levels; l++) {
            u32* rec = P.code.data() + (size_t)l * REC_WORDS * LANES;
            const Level& L = levels[l];
            if (L.special >= 0) {
                const Node& s = nodes[L.special];
                if ((int)s.deps.size() > 32) throw std::runtime_error("pvm: too many special arguments");
                for (int gl = 0; gl < LANES; gl++) {
                    rec[0 * LANES + gl] = gl < (int)s.deps.size() ? (u32)nodes[s.deps[gl]].slot : 0;
                    rec[1 * LANES + gl] = (u32)s.special_id << 8;
                    for (int t = 0; t < NMAX; t++) rec[(2 + t) * LANES + gl] = zero_term;
                }
                continue;
            }
            if (L.exp_id >= 0) {
                for (int gl = 0; gl < LANES; gl++) {
                    const int i = gl < EXP_LANES ? L.lane_node[gl] : -1;
                    for (int t = 0; t < NMAX; t++) rec[(2 + t) * LANES + gl] = zero_term;
                    rec[0 * LANES + gl] = i >= 0 ? ((u32)nodes[i].slot | W0_STORE) : 0;
                    if (i >= 0) rec[2 * LANES + gl] = (u32)nodes[nodes[i].deps[0]].slot | ((u32)(scratch_base + gl * EXP_TABLE) << 12);
                    rec[1 * LANES + gl] = ((u32)SP_EXP << 8) | ((u32)L.exp_id << 16);
                }
                continue;
            }
            for (int w = 0; w < NW; w++) {
                int N = 0, nsub = 1; bool comb = false, comb4 = false;
                for (int lane = 0; lane < 32; lane++) {
                    const int i = L.lane_node[w * 32 + lane];
                    if (i < 0) continue;
                    N = std::max(N, nodes[i].n); nsub = std::max(nsub, nodes[i].nsub); comb |= nodes[i].split > 1; comb4 |= nodes[i].split == 4;
                }
                for (int lane = 0; lane < 32; lane++) {
                    const int gl = w * 32 + lane;
                    const int i = L.lane_node[gl];
                    u32 w0 = 0;
                    for (int t = 0; t < NMAX; t++) rec[(2 + t) * LANES + gl] = zero_term;
                    if (i >= 0) {
                        const Node& n = nodes[i];
                        const int half = L.lane_half[gl];
                        if (half == 0) w0 = (u32)n.slot | W0_STORE | (n.split > 1 ? W0_COMBINE : 0) | (n.split == 4 ? W0_COMBINE4 : 0);
                        if (half == 1 && n.split == 4) w0 = W0_COMBINE;   // lane l+8 collects lane l+24 first
                        for (int t = 0; t < n.n; t++) {
                            const Term& tm = n.terms[half * n.n + t];
                            const int c = tm.coef < 0 ? -tm.coef : tm.coef;
                            const u32 sh = c == 4 ? 2 : c == 2 ? 1 : 0;
                            rec[(2 + t) * LANES + gl] = (u32)nodes[tm.a].slot | ((u32)nodes[tm.b].slot << 12) | (tm.coef < 0 ? 1u << 24 : 0) | (sh << 25);
                        }
                    }
                    rec[0 * LANES + gl] = w0;
                    rec[1 * LANES + gl] = (u32)N | ((u32)nsub << 4) | (comb ? 1u << 6 : 0) | (comb4 ? 1u << 7 : 0);
                }
            }
        }
        return P;
    }
};

}  // namespace pvm
}  // namespace zk
