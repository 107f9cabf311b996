// Fq6 / Fq12 tower and the optimal-ate pairing on BN254 for Groth16 verification on the device
// (one proof per thread).  Replaces ark-groth16 0.5.0 `prepare_verifying_key` + `verify_proof`
// as called from rln/src/protocol/proof.rs:856-894 (Cargo.lock:172; ark-ec pairing engine).
//
// Tower: Fq2 = Fq[u]/(u²+1), Fq6 = Fq2[v]/(v³−ξ), Fq12 = Fq6[w]/(w²−v), ξ = 9+u.
// The running point of the Miller loop stays affine on the twist E'(Fq2): y² = x³ + 3/ξ; a line
// through R with slope λ evaluated at P ∈ G1 is  −yP + (λ·xP)·w + (yR − λ·xR)·w³.
#pragma once
#include "curve.cuh"

namespace zk {

struct Fq6 {
    Fq2 c0, c1, c2;
    static HD Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static HD Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    HD bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    HD Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    HD Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    HD Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    HDN Fq6 operator*(const Fq6& o) const {
        Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
        Fq2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi() + t0;
        Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
        Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
        return {r0, r1, r2};
    }
    HD Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
    HD Fq6 scale2(const Fq2& k) const { return {c0 * k, c1 * k, c2 * k}; }           // × an Fq2 scalar
    HD Fq6 scale1(const Fq& k) const { return {c0.scale(k), c1.scale(k), c2.scale(k)}; }   // × an Fq scalar
    // × (d0 + d1·v): 5 Fq2 products instead of 6 + the ξ folds
    HDN Fq6 mul_by_01(const Fq2& d0, const Fq2& d1) const {
        const Fq2 a = c0 * d0, b = c1 * d1;
        const Fq2 r0 = ((c1 + c2) * d1 - b).mul_xi() + a;
        const Fq2 r1 = (c0 + c1) * (d0 + d1) - a - b;
        const Fq2 r2 = (c0 + c2) * d0 - a + b;
        return {r0, r1, r2};
    }
    HDN Fq6 inv() const {
        Fq2 A = c0.sqr() - (c1 * c2).mul_xi();
        Fq2 B = c2.sqr().mul_xi() - c0 * c1;
        Fq2 C = c1.sqr() - c0 * c2;
        Fq2 Fi = ((c2 * B + c1 * C).mul_xi() + c0 * A).inv();
        return {A * Fi, B * Fi, C * Fi};
    }
};

struct Fq12 {
    Fq6 c0, c1;
    static HD Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    HD bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    HDN Fq12 operator*(const Fq12& o) const {
        Fq6 t0 = c0 * o.c0, t1 = c1 * o.c1;
        Fq6 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1;
        return {t0 + t1.mul_v(), r1};
    }
    HDN Fq12 sqr() const {  // complex squaring
        Fq6 ab = c0 * c1;
        Fq6 t = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
        return {t, ab + ab};
    }
    HD Fq12 conj() const { return {c0, c1.neg()}; }
    // × (l0 + l3·w + l4·w³) with l0, l3, l4 ∈ Fq2 — the shape of every line of the Miller loop on a D-type twist
    // (ark-ff Fp12::mul_by_034): 3 + 5 + 5 Fq2 products instead of 18
    HDN Fq12 mul_by_034(const Fq2& l0, const Fq2& l3, const Fq2& l4) const {
        const Fq6 a = c0.scale2(l0);
        const Fq6 b = c1.mul_by_01(l3, l4);
        const Fq6 e = (c0 + c1).mul_by_01(l0 + l3, l4);
        return {a + b.mul_v(), e - a - b};
    }
    // the same with l0 ∈ Fq (lines of a FIXED G2 point, evaluated as −y_P + (λ·x_P)·w + c·w³): the first block is 6 Fq products
    HDN Fq12 mul_by_line_fq(const Fq& l0, const Fq2& l3, const Fq2& l4) const {
        const Fq6 a = c0.scale1(l0);
        const Fq6 b = c1.mul_by_01(l3, l4);
        const Fq6 e = (c0 + c1).mul_by_01(Fq2{l0 + l3.a, l3.b}, l4);
        return {a + b.mul_v(), e - a - b};
    }
    // squaring in the cyclotomic subgroup (elements of norm 1 over Fq6, i.e. everything after the easy part of the final
    // exponentiation): Granger–Scott, "Faster squaring in the cyclotomic subgroup of sixth degree extensions" — three Fq4
    // squarings (2 Fq2 squarings + 1 each) instead of a full Fq12 squaring
    HDN Fq12 cyclotomic_sqr() const {
        // Fq4 squaring (a + b·s)², s² = ξ: returns (a² + ξ·b², 2ab)
        auto fp4_sqr = [](const Fq2& a, const Fq2& b, Fq2& r0, Fq2& r1) {
            const Fq2 t0 = a.sqr(), t1 = b.sqr();
            r0 = t1.mul_xi() + t0;
            r1 = (a + b).sqr() - t0 - t1;
        };
        const Fq2 &z0 = c0.c0, &z4 = c0.c1, &z3 = c0.c2, &z2 = c1.c0, &z1 = c1.c1, &z5 = c1.c2;
        Fq2 t0, t1, t2, t3, t4, t5;
        fp4_sqr(z0, z1, t0, t1);
        fp4_sqr(z2, z3, t2, t3);
        fp4_sqr(z4, z5, t4, t5);
        Fq12 r;
        r.c0.c0 = (t0 - z0).dbl() + t0;
        r.c1.c1 = (t1 + z1).dbl() + t1;
        const Fq2 t5x = t5.mul_xi();
        r.c1.c0 = (t5x + z2).dbl() + t5x;
        r.c0.c2 = (t4 - z3).dbl() + t4;
        r.c0.c1 = (t2 - z4).dbl() + t2;
        r.c1.c2 = (t3 + z5).dbl() + t3;
        return r;
    }
    HDN Fq12 inv() const {
        Fq6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * d, (c1 * d).neg()};
    }
};

// Constants that depend only on q (computed once on the host, see pairing_tables_init in host code)
struct PairingTables {
    Fq2 gamma2;    // ξ^((q−1)/3)
    Fq2 gamma3;    // ξ^((q−1)/2)
    Fq2 twist_b;   // 3/ξ: the constant of the twist E'(Fq2): y² = x³ + 3/(9+u)
    Fq2 frob1[6];  // ξ^(k(q−1)/6), k = 0..5: coefficients of the q-power Frobenius on Fq12
    Fq frob2[6];   // ξ^(k(q²−1)/6), k = 0..5 (lie in Fq)
    u32 hard[24];  // (q⁴ − q² + 1)/r, little-endian words (761 bits) — reference value for the generic path (tests)
};

HD Fq12 line_eval(const Fq2& lam, const G2Affine& R, const G1Affine& P) {
    Fq12 l;
    l.c0 = {Fq2{P.y.neg(), Fq::zero()}, Fq2::zero(), Fq2::zero()};
    l.c1 = {lam.scale(P.x), R.y - lam * R.x, Fq2::zero()};
    return l;
}

// Miller loop f_{6x+2,Q}(P)·(Frobenius correction lines); no final exponentiation.
HDN Fq12 miller_loop(const PairingTables* pt, const G2Affine& Qp, const G1Affine& P) {
    if (Qp.is_inf() || P.is_inf()) return Fq12::one();
    const u64 ATE_LOW = 0x9d797039be763ba8ULL;  // 6x+2 = 2^64 + ATE_LOW
    G2Affine R = Qp;
    Fq12 f = Fq12::one();
    for (int i = 63; i >= 0; i--) {
        {  // doubling step
            Fq2 xx = R.x.sqr();
            Fq2 lam = (xx.dbl() + xx) * R.y.dbl().inv();
            f = f.sqr() * line_eval(lam, R, P);
            Fq2 x3 = lam.sqr() - R.x.dbl();
            R.y = lam * (R.x - x3) - R.y;
            R.x = x3;
        }
        if ((ATE_LOW >> i) & 1) {
            Fq2 lam = (Qp.y - R.y) * (Qp.x - R.x).inv();
            f = f * line_eval(lam, R, P);
            Fq2 x3 = lam.sqr() - R.x - Qp.x;
            R.y = lam * (R.x - x3) - R.y;
            R.x = x3;
        }
    }
    G2Affine Q1 = {Qp.x.conj() * pt->gamma2, Qp.y.conj() * pt->gamma3};
    G2Affine Q2 = {Q1.x.conj() * pt->gamma2, (Q1.y.conj() * pt->gamma3).neg()};
    for (int k = 0; k < 2; k++) {
        const G2Affine& S = k == 0 ? Q1 : Q2;
        Fq2 lam = (S.y - R.y) * (S.x - R.x).inv();
        f = f * line_eval(lam, R, P);
        Fq2 x3 = lam.sqr() - R.x - S.x;
        R.y = lam * (R.x - x3) - R.y;
        R.x = x3;
    }
    return f;
}

// ---- inversion-free Miller loop ------------------------------------------------------------------------------------------------
// The running point stays in homogeneous projective coordinates on the twist, so a step costs ≈ 25 Fq products of point arithmetic
// instead of an Fq2 inversion (≈ 400); lines come out as (l0, l3, l4) for Fq12::mul_by_034 after scaling by y_P and x_P.  Formulas:
// Costello–Lange–Naehrig "Faster explicit formulas for computing pairings over ordinary curves", as arranged in ark-ec 0.5.0
// models/bn/g2.rs (doubling_step / addition_step, TwistType::D) — the code path ark-groth16's verifier runs for BN254.
struct G2Proj { Fq2 x, y, z; };
HD Fq halve(const Fq& a) {   // a/2 on the Montgomery residue: halving is linear
    u32 t[8];
    u32 carry = 0;
    if (a.l[0] & 1) {
        u32 q[8];
        for (int i = 0; i < 8; i++) q[i] = FqCfg::p(i);
        carry = Fq::raw_add(t, a.l, q);
    } else {
        for (int i = 0; i < 8; i++) t[i] = a.l[i];
    }
    Fq r;
    for (int i = 0; i < 7; i++) r.l[i] = (t[i] >> 1) | (t[i + 1] << 31);
    r.l[7] = (t[7] >> 1) | (carry << 31);
    return r;
}
HD Fq2 halve(const Fq2& a) { return {halve(a.a), halve(a.b)}; }
// R ← 2R; line through R (tangent): l0 = −h (× y_P), l3 = 3x² (× x_P), l4 = e − b
HD void proj_double(const PairingTables* pt, G2Proj& r, Fq2& l0, Fq2& l3, Fq2& l4) {
    const Fq2 a = halve(r.x * r.y);
    const Fq2 b = r.y.sqr();
    const Fq2 c = r.z.sqr();
    const Fq2 e = pt->twist_b * (c.dbl() + c);
    const Fq2 f = e.dbl() + e;
    const Fq2 g = halve(b + f);
    const Fq2 h = (r.y + r.z).sqr() - (b + c);
    const Fq2 i = e - b;
    const Fq2 j = r.x.sqr();
    const Fq2 e2 = e.sqr();
    r.x = a * (b - f);
    r.y = g.sqr() - (e2.dbl() + e2);
    r.z = b * h;
    l0 = h.neg();
    l3 = j.dbl() + j;
    l4 = i;
}
// R ← R + Q (Q affine); line through R and Q: l0 = λ (× y_P), l3 = −θ (× x_P), l4 = θ·x_Q − λ·y_Q
HD void proj_add(G2Proj& r, const G2Affine& q, Fq2& l0, Fq2& l3, Fq2& l4) {
    const Fq2 theta = r.y - q.y * r.z;
    const Fq2 lambda = r.x - q.x * r.z;
    const Fq2 c = theta.sqr();
    const Fq2 d = lambda.sqr();
    const Fq2 e = lambda * d;
    const Fq2 f = r.z * c;
    const Fq2 g = r.x * d;
    const Fq2 h = e + f - g.dbl();
    r.x = lambda * h;
    r.y = theta * (g - h) - e * r.y;
    r.z = r.z * e;
    l0 = lambda;
    l3 = theta.neg();
    l4 = theta * q.x - lambda * q.y;
}
// The three pairings of a Groth16 check that depend on the proof, in ONE loop with one squaring per bit:
//   f = Miller(−A, B) · Miller(vk_x, γ) · Miller(C, δ)     (B varies → projective steps; γ, δ fixed → precomputed affine lines)
// Neither A / vk_x / C at infinity nor B at infinity contribute a factor (e(O, ·) = e(·, O) = 1).
HDN Fq12 miller_loop_groth16(const PairingTables* pt, const G2Affine& Bq, const G1Affine& negA, const Fq2* __restrict__ g_lam,
                             const Fq2* __restrict__ g_c, const G1Affine& vkx, const Fq2* __restrict__ d_lam, const Fq2* __restrict__ d_c,
                             const G1Affine& C) {
    const u64 ATE_LOW = 0x9d797039be763ba8ULL;  // 6x+2 = 2^64 + ATE_LOW
    const bool var_on = !(Bq.is_inf() || negA.is_inf()), g_on = !vkx.is_inf(), d_on = !C.is_inf();
    G2Proj R = {Bq.x, Bq.y, Fq2::one()};
    const Fq g_ny = vkx.y.neg(), d_ny = C.y.neg();
    Fq12 f = Fq12::one();
    int n = 0;
    Fq2 l0, l3, l4;
    auto fixed_lines = [&](int k) {
        if (g_on) f = f.mul_by_line_fq(g_ny, g_lam[k].scale(vkx.x), g_c[k]);
        if (d_on) f = f.mul_by_line_fq(d_ny, d_lam[k].scale(C.x), d_c[k]);
    };
    for (int i = 63; i >= 0; i--) {
        if (i != 63) f = f.sqr();   // f = 1 before the first step
        if (var_on) {
            proj_double(pt, R, l0, l3, l4);
            f = f.mul_by_034(l0.scale(negA.y), l3.scale(negA.x), l4);
        }
        fixed_lines(n++);
        if ((ATE_LOW >> i) & 1) {
            if (var_on) {
                proj_add(R, Bq, l0, l3, l4);
                f = f.mul_by_034(l0.scale(negA.y), l3.scale(negA.x), l4);
            }
            fixed_lines(n++);
        }
    }
    const G2Affine Q1 = {Bq.x.conj() * pt->gamma2, Bq.y.conj() * pt->gamma3};
    const G2Affine Q2 = {Q1.x.conj() * pt->gamma2, (Q1.y.conj() * pt->gamma3).neg()};
    for (int k = 0; k < 2; k++) {
        if (var_on) {
            proj_add(R, k == 0 ? Q1 : Q2, l0, l3, l4);
            f = f.mul_by_034(l0.scale(negA.y), l3.scale(negA.x), l4);
        }
        fixed_lines(n++);
    }
    return f;
}
// single variable pairing with the projective steps (tests: must agree with miller_loop after the final exponentiation)
HDN Fq12 miller_loop_proj(const PairingTables* pt, const G2Affine& Qp, const G1Affine& P) {
    if (Qp.is_inf() || P.is_inf()) return Fq12::one();
    const u64 ATE_LOW = 0x9d797039be763ba8ULL;
    G2Proj R = {Qp.x, Qp.y, Fq2::one()};
    Fq12 f = Fq12::one();
    Fq2 l0, l3, l4;
    for (int i = 63; i >= 0; i--) {
        if (i != 63) f = f.sqr();
        proj_double(pt, R, l0, l3, l4);
        f = f.mul_by_034(l0.scale(P.y), l3.scale(P.x), l4);
        if ((ATE_LOW >> i) & 1) {
            proj_add(R, Qp, l0, l3, l4);
            f = f.mul_by_034(l0.scale(P.y), l3.scale(P.x), l4);
        }
    }
    const G2Affine Q1 = {Qp.x.conj() * pt->gamma2, Qp.y.conj() * pt->gamma3};
    const G2Affine Q2 = {Q1.x.conj() * pt->gamma2, (Q1.y.conj() * pt->gamma3).neg()};
    proj_add(R, Q1, l0, l3, l4);
    f = f.mul_by_034(l0.scale(P.y), l3.scale(P.x), l4);
    proj_add(R, Q2, l0, l3, l4);
    f = f.mul_by_034(l0.scale(P.y), l3.scale(P.x), l4);
    return f;
}

// G2 membership of a point already known to be on the twist: ψ(P) = [6x²]P, the criterion ark-bn254 0.5.0 applies when it
// deserialises a G2 point (g2.rs, is_in_correct_subgroup_assuming_on_curve; on the r-torsion ψ acts as q ≡ 6x² mod r) —
// a 127-bit multiple instead of the 254-bit [r]P.  Kept as the cross-check of the faster test below.
HDN bool g2_in_subgroup_6x2(const PairingTables* pt, const G2Affine& p) {
    const u32 six_x2[8] = {0xe87cfd46u, 0xf83e9682u, 0xeeb859fbu, 0x6f4d8248u, 0, 0, 0, 0};
    const G2Affine lhs = G2XYZZ::from_affine(p).mul(six_x2).to_affine();
    const G2Affine psi = {p.x.conj() * pt->gamma2, p.y.conj() * pt->gamma3};
    return !lhs.is_inf() && lhs.x == psi.x && lhs.y == psi.y;
}
// The same membership decided with ONE 63-bit multiple and no inversion:
//     [r]P = O  ⇔  [x+1]P + ψ([x]P) + ψ²([x]P) = ψ³([2x]P)        (x = 4965661367192848881, the BN parameter)
// (q ≡ 6x², q² … expressed through ψ on the r-torsion; the identity gnark-crypto's bn254 G2 IsInSubGroup uses.)  Both tests
// are exact, so they accept the same points; tests/test_host_emul.py compares them on subgroup and non-subgroup twist points.
HD G2XYZZ g2_psi(const PairingTables* pt, const G2XYZZ& p) {
    return {p.X.conj() * pt->gamma2, p.Y.conj() * pt->gamma3, p.ZZ.conj(), p.ZZZ.conj()};
}
HDN bool g2_in_subgroup(const PairingTables* pt, const G2Affine& p) {
    if (p.is_inf()) return false;
    const u64 X = 4965661367192848881ULL;
    const u32 xs[8] = {(u32)X, (u32)(X >> 32), 0, 0, 0, 0, 0, 0};
    const G2XYZZ P = G2XYZZ::from_affine(p);
    G2XYZZ xP = G2XYZZ::infinity();
    for (int i = 62; i >= 0; i--) {   // bit 62 is the leading one
        xP = xP.dbl();
        if ((xs[i >> 5] >> (i & 31)) & 1) xP.add_affine(p);
    }
    G2XYZZ lhs = xP;
    lhs.add(P);                                   // [x+1]P
    const G2XYZZ psi1 = g2_psi(pt, xP), psi2 = g2_psi(pt, psi1);
    lhs.add(psi1);
    lhs.add(psi2);
    const G2XYZZ rhs = g2_psi(pt, g2_psi(pt, g2_psi(pt, xP.dbl())));
    if (lhs.is_inf() || rhs.is_inf()) return lhs.is_inf() && rhs.is_inf();
    return lhs.X * rhs.ZZ == rhs.X * lhs.ZZ && lhs.Y * rhs.ZZZ == rhs.Y * lhs.ZZZ;
}

// Line coefficients of the Miller loop for a FIXED G2 point (γ₂, δ₂ of the verifying key): per step the slope λ and
// c = y_R − λ·x_R, so evaluating at P needs no point arithmetic and no inversion.
constexpr int MILLER_STEPS = 64 + 36 + 2;  // 64 doublings, popcount(ATE_LOW) = 36 additions, 2 Frobenius corrections
struct FixedLines {
    Fq2 lam[MILLER_STEPS];
    Fq2 c[MILLER_STEPS];
};
HDN void precompute_lines(const PairingTables* pt, const G2Affine& Qp, FixedLines& out) {
    const u64 ATE_LOW = 0x9d797039be763ba8ULL;
    G2Affine R = Qp;
    int n = 0;
    auto step = [&](const Fq2& lam, const G2Affine* S) {
        out.lam[n] = lam;
        out.c[n] = R.y - lam * R.x;
        n++;
        Fq2 x3 = S ? lam.sqr() - R.x - S->x : lam.sqr() - R.x.dbl();
        R.y = lam * (R.x - x3) - R.y;
        R.x = x3;
    };
    for (int i = 63; i >= 0; i--) {
        Fq2 xx = R.x.sqr();
        step((xx.dbl() + xx) * R.y.dbl().inv(), nullptr);
        if ((ATE_LOW >> i) & 1) step((Qp.y - R.y) * (Qp.x - R.x).inv(), &Qp);
    }
    G2Affine Q1 = {Qp.x.conj() * pt->gamma2, Qp.y.conj() * pt->gamma3};
    G2Affine Q2 = {Q1.x.conj() * pt->gamma2, (Q1.y.conj() * pt->gamma3).neg()};
    step((Q1.y - R.y) * (Q1.x - R.x).inv(), &Q1);
    step((Q2.y - R.y) * (Q2.x - R.x).inv(), &Q2);
}
HDN Fq12 miller_loop_fixed(const Fq2* __restrict__ lam, const Fq2* __restrict__ cc, const G1Affine& P) {
    if (P.is_inf()) return Fq12::one();
    const u64 ATE_LOW = 0x9d797039be763ba8ULL;
    Fq12 f = Fq12::one();
    int n = 0;
    const Fq ny = P.y.neg();
    auto line = [&](int k) {
        Fq12 l;
        l.c0 = {Fq2{ny, Fq::zero()}, Fq2::zero(), Fq2::zero()};
        l.c1 = {lam[k].scale(P.x), cc[k], Fq2::zero()};
        return l;
    };
    for (int i = 63; i >= 0; i--) {
        f = f.sqr() * line(n++);
        if ((ATE_LOW >> i) & 1) f = f * line(n++);
    }
    f = f * line(n++);
    f = f * line(n++);
    return f;
}

HDN Fq12 frobenius2(const PairingTables* pt, const Fq12& f) {
    Fq12 r;  // coefficient of w^k is scaled by frob2[k]; c0 = (w⁰,w²,w⁴), c1 = (w¹,w³,w⁵)
    r.c0 = {f.c0.c0, f.c0.c1.scale(pt->frob2[2]), f.c0.c2.scale(pt->frob2[4])};
    r.c1 = {f.c1.c0.scale(pt->frob2[1]), f.c1.c1.scale(pt->frob2[3]), f.c1.c2.scale(pt->frob2[5])};
    return r;
}

// q-power Frobenius: (Σ a_k w^k)^q = Σ conj(a_k)·ξ^{k(q−1)/6}·w^k
HDN Fq12 frobenius1(const PairingTables* pt, const Fq12& f) {
    Fq12 r;
    r.c0 = {f.c0.c0.conj(), f.c0.c1.conj() * pt->frob1[2], f.c0.c2.conj() * pt->frob1[4]};
    r.c1 = {f.c1.c0.conj() * pt->frob1[1], f.c1.c1.conj() * pt->frob1[3], f.c1.c2.conj() * pt->frob1[5]};
    return r;
}
// f^u for the BN parameter u = 4965661367192848881 (63 bits)
HDN Fq12 pow_u(const Fq12& f) {
    const u64 U = 4965661367192848881ULL;
    Fq12 r = f;
    for (int i = 61; i >= 0; i--) {  // bit 62 is the leading one
        r = r.cyclotomic_sqr();      // only ever called on elements of the cyclotomic subgroup (after the easy part)
        if ((U >> i) & 1) r = r * f;
    }
    return r;
}
// generic hard part (plain exponentiation by (q⁴−q²+1)/r); kept as the cross-check of the addition chain below
HDN Fq12 final_exponentiation_generic(const PairingTables* pt, const Fq12& f) {
    Fq12 t = f.conj() * f.inv();      // ^(q⁶−1)
    t = frobenius2(pt, t) * t;        // ^(q²+1)
    Fq12 r = Fq12::one();             // ^((q⁴−q²+1)/r)
    for (int i = 24 * 32 - 1; i >= 0; i--) {
        r = r.sqr();
        if ((pt->hard[i >> 5] >> (i & 31)) & 1) r = r * t;
    }
    return r;
}
// Final exponentiation with the BN hard-part addition chain (Scott et al., "On the final exponentiation for
// calculating pairings on ordinary elliptic curves"): three powers of u plus Frobenius maps.  After the easy part
// the element is unitary, so inversion is conjugation.  The result is f^{(q¹²−1)/r·c} for a fixed c coprime to r —
// equal to 1 exactly when the generic exponentiation gives 1, which is all a verifier needs.
HDN Fq12 final_exponentiation(const PairingTables* pt, const Fq12& f) {
    Fq12 t1 = f.conj() * f.inv();       // ^(q⁶−1)
    t1 = frobenius2(pt, t1) * t1;       // ^(q²+1)
    Fq12 fp = frobenius1(pt, t1), fp2 = frobenius2(pt, t1), fp3 = frobenius1(pt, fp2);
    Fq12 fu = pow_u(t1), fu2 = pow_u(fu), fu3 = pow_u(fu2);
    Fq12 y3 = frobenius1(pt, fu), fu2p = frobenius1(pt, fu2), fu3p = frobenius1(pt, fu3), y2 = frobenius2(pt, fu2);
    Fq12 y0 = fp * fp2 * fp3;
    Fq12 y1 = t1.conj(), y5 = fu2.conj();
    y3 = y3.conj();
    Fq12 y4 = (fu * fu2p).conj();
    Fq12 y6 = (fu3 * fu3p).conj();
    Fq12 t0 = y6.cyclotomic_sqr() * y4 * y5;
    Fq12 t2 = y3 * y5 * t0;
    t0 = t0 * y2;
    t2 = (t2.cyclotomic_sqr() * t0).cyclotomic_sqr();
    t0 = t2 * y1;
    t2 = t2 * y0;
    return t0.cyclotomic_sqr() * t2;
}

}  // namespace zk
