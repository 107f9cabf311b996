// The data-dependent steps of the pairing VM (verify_vm.hpp), as host+device code shared by k_verify_vm and the host interpreter
// of tests/host_emul: parsing the compressed proof (ark-serialize flags, rln/src/protocol/proof.rs:456-470), choosing roots and
// signs after the square-root chains, the final comparisons.  Every decision follows k_verify.cu's decompress_g1 / decompress_g2 /
// k_verify, which remains the reference for anything reported as FALLBACK.
#pragma once
#include "verify_vm.hpp"

namespace zk {
namespace pvm {

// flags kept between the prologue and SP_SELECT: bit 0 / 1 / 2 = "y is the larger root" bit of A / B / C
struct ProofFlags { u32 larger; };

HD void pv_load_words(const uint8_t* p, u32* w) {
    for (int i = 0; i < 8; i++) w[i] = (u32)p[4 * i] | ((u32)p[4 * i + 1] << 8) | ((u32)p[4 * i + 2] << 16) | ((u32)p[4 * i + 3] << 24);
}
HD bool pv_canonical(const u32* w) {
    u32 q[8];
    for (int i = 0; i < 8; i++) q[i] = FqCfg::p(i);
    return Fq::raw_cmp(w, q) < 0;
}
// canonical value c: is c > q − c ?
HD bool pv_larger_half(const Fq& canon) {
    u32 q[8], n[8];
    for (int i = 0; i < 8; i++) q[i] = FqCfg::p(i);
    Fq::raw_sub(n, q, canon.l);
    return Fq::raw_cmp(canon.l, n) > 0;
}
// proof = A (32 B) | B (64 B) | C (32 B), ark-compressed.  Writes the raw x coordinates into their input slots.
HD u32 pv_prologue(const uint8_t* proof, Fq* slots, ProofFlags& fl) {
    u32 xa[8], xc[8], xb0[8], xb1[8];
    pv_load_words(proof, xa);
    pv_load_words(proof + 32, xb0);
    pv_load_words(proof + 64, xb1);
    pv_load_words(proof + 96, xc);
    const u32 fa = xa[7] >> 30, fb = xb1[7] >> 30, fc = xc[7] >> 30;
    xa[7] &= 0x3fffffffu; xb1[7] &= 0x3fffffffu; xc[7] &= 0x3fffffffu;
    if (fa == 3 || fb == 3 || fc == 3) return ST_MALFORMED;
    if (!pv_canonical(xa) || !pv_canonical(xb0) || !pv_canonical(xb1) || !pv_canonical(xc)) return ST_MALFORMED;
    if ((fa | fb | fc) & 1) return ST_FALLBACK;   // a point at infinity: the complete formulas of k_verify decide
    fl.larger = ((fa >> 1) & 1) | (((fb >> 1) & 1) << 1) | (((fc >> 1) & 1) << 2);
    for (int i = 0; i < 8; i++) { slots[S_XA].l[i] = xa[i]; slots[S_XC].l[i] = xc[i]; slots[S_XB0].l[i] = xb0[i]; slots[S_XB1].l[i] = xb1[i]; }
    return ST_RUNNING;
}

// SP_EXP: a^e with 4-bit windows; the table a¹…a¹⁵ goes to the lane's scratch slots tab[1..15]
HDN Fq pv_pow(const Fq& a, const u32* e, Fq* tab) {
    tab[1] = a;
    for (int k = 2; k < 16; k++) tab[k] = (k & 1) ? tab[k - 1] * a : tab[k / 2].sqr();
    int top = 63;
    while (top > 0 && ((e[top >> 3] >> ((top & 7) * 4)) & 15) == 0) top--;
    Fq r = tab[(e[top >> 3] >> ((top & 7) * 4)) & 15];
    for (int d = top - 1; d >= 0; d--) {
        r = r.sqr().sqr().sqr().sqr();
        const u32 dg = (e[d >> 3] >> ((d & 7) * 4)) & 15;
        if (dg) r = r * tab[dg];
    }
    return r;
}

// SP_SELECT arguments (slot numbers), in order
enum SelectArg {
    SA_CHK_A = 0, SA_RHS_A, SA_Y_A, SA_YC_A,          // y², x³+3, y, canonical y of A
    SA_CHK_C, SA_RHS_C, SA_Y_C, SA_YC_C,              // … of C
    SA_CHK_ALPHA, SA_NORM,                            // α² and N(x³+b') of B
    SA_D_P, SA_CHK_P, SA_X0_P, SA_X1_P, SA_X0C_P, SA_X1C_P,   // candidate δ = (a + α)/2: δ, x0², x0, x1, canonical x0, x1
    SA_D_M, SA_CHK_M, SA_X0_M, SA_X1_M, SA_X0C_M, SA_X1C_M,   // candidate δ = (a − α)/2
    SA_RHS_B_IM, SA_VKX_DEP,
    SA_OUT_NAY, SA_OUT_CY, SA_OUT_BY0, SA_OUT_BY1, SA_COUNT
};
// returns the new status (ST_RUNNING to go on)
HD u32 pv_select(Fq* slots, const u32* arg, const ProofFlags& fl) {
    auto S = [&](int a) -> Fq& { return slots[arg[a]]; };
    // G1 points: y = rhs^((q+1)/4) must square to rhs (decompress_g1)
    if (S(SA_CHK_A) != S(SA_RHS_A) || S(SA_CHK_C) != S(SA_RHS_C)) return ST_MALFORMED;
    {
        Fq y = S(SA_Y_A);
        if (pv_larger_half(S(SA_YC_A)) != ((fl.larger & 1) != 0)) y = y.neg();
        S(SA_OUT_NAY) = y.neg();   // the Miller loop takes −A
        y = S(SA_Y_C);
        if (pv_larger_half(S(SA_YC_C)) != ((fl.larger & 4) != 0)) y = y.neg();
        S(SA_OUT_CY) = y;
    }
    // G2 point (fq2_sqrt, complex method): a + b·u with b = 0 takes another route there
    if (S(SA_RHS_B_IM).is_zero()) return ST_FALLBACK;
    if (S(SA_CHK_ALPHA) != S(SA_NORM)) return ST_MALFORMED;
    int base;
    if (S(SA_CHK_P) == S(SA_D_P)) base = SA_D_P;
    else if (S(SA_CHK_M) == S(SA_D_M)) base = SA_D_M;
    else return ST_MALFORMED;
    if (slots[arg[base + 2]].is_zero()) return ST_FALLBACK;   // x0 = 0: no inverse
    Fq y0 = slots[arg[base + 2]], y1 = slots[arg[base + 3]];
    const Fq c0 = slots[arg[base + 4]], c1 = slots[arg[base + 5]];
    const bool larger = y1.is_zero() ? pv_larger_half(c0) : pv_larger_half(c1);
    if (larger != ((fl.larger & 2) != 0)) { y0 = y0.neg(); y1 = y1.neg(); }
    S(SA_OUT_BY0) = y0;
    S(SA_OUT_BY1) = y1;
    return ST_RUNNING;
}

// SP_FINAL arguments: the 12 coefficients of the final exponentiation's result (w⁰.re, w⁰.im, w¹.re, …), the cross products
// of the membership identity  lhs.X·rhs.ZZ, rhs.X·lhs.ZZ, lhs.Y·rhs.ZZZ, rhs.Y·lhs.ZZZ  (re, im each), then lhs.ZZ, rhs.ZZ
enum FinalArg { FA_F = 0, FA_CROSS = 12, FA_ZZ = 20, FA_COUNT = 24 };
HD u32 pv_final(const Fq* slots, const u32* arg) {
    auto S = [&](int a) -> const Fq& { return slots[arg[a]]; };
    // an exceptional addition anywhere in the membership chain leaves ZZ = 0 from there on
    if ((S(FA_ZZ).is_zero() && S(FA_ZZ + 1).is_zero()) || (S(FA_ZZ + 2).is_zero() && S(FA_ZZ + 3).is_zero())) return ST_FALLBACK;
    for (int k = 0; k < 4; k += 2)
        if (S(FA_CROSS + 2 * k) != S(FA_CROSS + 2 * k + 2) || S(FA_CROSS + 2 * k + 1) != S(FA_CROSS + 2 * k + 3)) return ST_MALFORMED;   // B ∉ G2
    bool one = S(FA_F) == Fq::one();
    for (int k = 1; k < 12; k++) one = one && S(FA_F + k).is_zero();
    return one ? ST_VALID : ST_INVALID;
}

}  // namespace pvm
}  // namespace zk
