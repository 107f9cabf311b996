// The Groth16 verification of k_verify.cu written once more, symbolically, for the pairing VM (verify_vm.hpp): every formula
// below follows the device code it replaces for the common case — decompress_g1 / decompress_g2 / g2_in_subgroup (k_verify.cu,
// tower.cuh), miller_loop_groth16 with ark-ec's projective steps, final_exponentiation — rearranged so that each value is a sum
// of products of values that already exist (linear operations cost nothing, they are folded into the next sum).  Host code:
// runs once per verifying key (ark-groth16 prepare_verifying_key is the closest thing in the reference,
// rln/src/protocol/proof.rs:869-871).
#pragma once
#include "verify_vm_special.cuh"

namespace zk {
namespace pvm {

struct VerifyKeyHost {
    const FixedLines* gamma;   // precompute_lines(γ₂)
    const FixedLines* delta;   // precompute_lines(δ₂)
    Fq12 ml_alpha_beta;        // miller_loop(β₂, α₁)
};

struct ProgramBuilder : Builder {
    typedef Builder B;
    const PairingTables& pt;
    Fq2 xi_;
    explicit ProgramBuilder(const PairingTables& t) : pt(t) { xi_ = {Fq::from_u32(9), Fq::from_u32(1)}; init_frob(t); }

    static void exp_words(u32* e, int add) {   // (q + add)/4 for add = +1 / −3; add = −2 → q − 2 (not divided)
        for (int i = 0; i < 8; i++) e[i] = FqCfg::p(i);
        if (add == -2) { e[0] -= 2; return; }
        if (add == 1) e[0] += 1; else e[0] -= 3;   // low word 0xd87cfd47: neither carries nor borrows
        for (int i = 0; i < 7; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
        e[7] >>= 2;
    }
    // ξ^p · s for p = 0..n−1, each materialised
    std::vector<S2> xi_powers(const S2& s, int n) {
        std::vector<S2> r;
        r.push_back(mat(s));
        for (int p = 1; p < n; p++) r.push_back(mat(mul_xi(r.back())));
        return r;
    }
    // a sparse line ℓ0 + ℓ1·w + ℓ3·w³ together with ξ^p multiples: P[p][k] (k ∈ {0,1,3})
    struct Line { S2 P[4][6]; };
    // M · line, with the result's ξ and ξ² multiples: out.c = Σ M_i·ℓ^(0 | 1 when wrapped), out.x = …^(1 | 2), out.x2 = …^(2 | 3)
    S12 mul_line(const S12& M, const Line& ln) {
        S12 r;
        for (int e = 0; e < 3; e++)
            for (int k = (e == 2 ? 1 : 0); k < 6; k++) {
                Acc2 acc(*this);
                for (int i = 0; i < 6; i++) {
                    int m = k - i, p = e;
                    if (m < 0) { m += 6; p++; }
                    acc.add(M.c[i], ln.P[p][m]);
                }
                (e == 0 ? r.c[k] : e == 1 ? r.x[k] : r.x2[k]) = acc.emit();
            }
        r.has_x = r.has_x2 = true;
        return r;
    }
    S12 mulx(S12 a, S12 b, bool want_x) {   // a·b with whichever operand already carries the multiples on the right
        auto fits = [&](const S12& t) { return t.has_x && (!want_x || t.has_x2); };
        if (!fits(b) && fits(a)) std::swap(a, b);
        if (want_x) ensure_x2(b); else ensure_x(b);
        return mul12(a, b, want_x);
    }
    S12 pow_u(S12 f) {   // tower.cuh pow_u: f^u, u = 4965661367192848881, f in the cyclotomic subgroup
        const u64 U = 4965661367192848881ULL;
        ensure_x2(f);
        S12 r = f;
        for (int i = 61; i >= 0; i--) {
            r = cyc_sqr(r);
            if ((U >> i) & 1) r = mul12(r, f, true);
        }
        return r;
    }

    Program build(const VerifyKeyHost& vk) {
        const Val XA = v(S_XA), XC = v(S_XC), R2 = v(S_R2), RAW1 = v(S_RAW1);
        const Val HALF = cfq(Fq::from_u32(2).inv());
        const int e_sqrt = EXP_SQRT, e_m3 = EXP_INV_SQRT, e_inv = EXP_INV;
        // ------------------------------------------------------------------ decompression (chains 1 and 2)
        const Val xA = B::mul(XA, R2), xC = B::mul(XC, R2);
        const S2 xB = {B::mul(v(S_XB0), R2), B::mul(v(S_XB1), R2)};
        auto g1_rhs = [&](const Val& x) { const Val x2 = B::mul(x, x); return dot({{x2, x}, {kconst(3), one()}}); };
        const Val rhsA = g1_rhs(xA), rhsC = g1_rhs(xC);
        const Val yA = pow_fixed(rhsA, e_sqrt), yC = pow_fixed(rhsC, e_sqrt);
        const S2 xB2 = B::mul(xB, xB);
        S2 rhsB;
        { Acc2 a(*this); a.add(xB2, xB); a.lin(cfq2(pt.twist_b)); rhsB = a.emit(); }
        const Val norm = dot({{rhsB.a, rhsB.a}, {rhsB.b, rhsB.b}});
        const Val alpha = pow_fixed(norm, e_sqrt);
        const Val bh = B::mul(rhsB.b, HALF);
        Val cand[2][6];
        for (int s = 0; s < 2; s++) {
            const Val d = B::mul(s == 0 ? rhsB.a + alpha : rhsB.a - alpha, HALF);
            const Val w = pow_fixed(d, e_m3);            // δ^((q−3)/4) = 1/√δ when δ is a square
            const Val x0 = B::mul(w, d), x1 = B::mul(bh, w);
            cand[s][0] = d; cand[s][1] = B::mul(x0, x0); cand[s][2] = x0; cand[s][3] = x1; cand[s][4] = B::mul(x0, RAW1); cand[s][5] = B::mul(x1, RAW1);
        }
        // vk_x arrives from its own warp
        const std::vector<Val> vkx = special(SP_VKX, {}, 4, {S_VX, S_VY, S_VZZ, S_VZZZ});
        std::vector<Val> sel_in = {B::mul(yA, yA), rhsA, yA, B::mul(yA, RAW1), B::mul(yC, yC), rhsC, yC, B::mul(yC, RAW1), B::mul(alpha, alpha), norm};
        for (int s = 0; s < 2; s++) for (int k = 0; k < 6; k++) sel_in.push_back(cand[s][k]);
        sel_in.push_back(rhsB.b);
        sel_in.push_back(vkx[2]);
        if ((int)sel_in.size() != SA_OUT_NAY) throw std::runtime_error("pvm: SP_SELECT layout");
        const std::vector<Val> sel = special(SP_SELECT, sel_in, 4);
        const Val nAy = sel[0], Cy = sel[1];
        const S2 yB = {sel[2], sel[3]};

        // ------------------------------------------------------------------ per-proof scalars of the lines
        const std::vector<S2> Y_p = xi_powers(S2{nAy, Val{}}, 3);      // (−A.y)·ξ^p
        const std::vector<S2> XA_p = xi_powers(S2{xA, Val{}}, 4);      // A.x·ξ^p
        const std::vector<S2> XQ_p = xi_powers(xB, 4), YQ_p = xi_powers(yB, 4);
        std::vector<S2> XI_p;                                           // ξ^p
        { Fq2 c = Fq2::one(); for (int p = 0; p < 4; p++) { XI_p.push_back(cfq2(c)); c = c * xi_; } }
        // fixed lines: ℓγ·(ZZ·ZZZ) = s0 + s1·λγ·w + s3·cγ·w³ for vk_x = (X/ZZ, Y/ZZZ);  ℓδ = t0 + t1·λδ·w + cδ·w³ for C
        const Val s0 = B::mul(vkx[1] * -1, vkx[2]), s1 = B::mul(vkx[0], vkx[3]), s3 = B::mul(vkx[2], vkx[3]);
        const Val c0 = Cy * -1, c1 = xC;   // ℓδ's scalars t0, t1
        const Val u00 = B::mul(s0, c0), u01 = B::mul(s0, c1), u10 = B::mul(s1, c0), u11 = B::mul(s1, c1), u30 = B::mul(s3, c0), u31 = B::mul(s3, c1);
        auto fixed_M = [&](int k) {
            const Fq2 lg = vk.gamma->lam[k], cg = vk.gamma->c[k], ld = vk.delta->lam[k], cd = vk.delta->c[k];
            const S2 LG = cfq2(lg), CG = cfq2(cg), LD = cfq2(ld), CD = cfq2(cd), LL = cfq2(lg * ld), XCC = cfq2(cg * cd * xi_), LGCD = cfq2(lg * cd), LDCG = cfq2(ld * cg);
            S12 M;
            { Acc2 a(*this); a.add_fq(XCC, s3); a.re.push_back({u00, one()}); M.c[0] = a.emit(); }
            { Acc2 a(*this); a.add_fq(LD, u01); a.add_fq(LG, u10); M.c[1] = a.emit(); }
            { Acc2 a(*this); a.add_fq(LL, u11); M.c[2] = a.emit(); }
            { Acc2 a(*this); a.add_fq(CD, s0); a.add_fq(CG, u30); M.c[3] = a.emit(); }
            { Acc2 a(*this); a.add_fq(LGCD, s1); a.add_fq(LDCG, u31); M.c[4] = a.emit(); }
            return M;
        };

        // ------------------------------------------------------------------ Miller loop
        const u64 ATE_LOW = 0x9d797039be763ba8ULL;   // 6x+2 = 2^64 + ATE_LOW
        struct RPt { S2 x, y, z, zt, zx, zy; } R;
        R.x = xB; R.y = yB; R.z = one2(); R.zt = cfq2(pt.twist_b + pt.twist_b + pt.twist_b); R.zx = xB; R.zy = yB;
        S12 f = one12();
        int n_fixed = 0;
        // the companions zt = 3b'·z, zx = x_B·z, zy = y_B·z travel with R so that no step needs a product of three values; the
        // ones nobody reads afterwards are dropped by the scheduler
        auto dbl_step = [&]() {   // R ← 2R (scaled by 4), returns the tangent line with its multiples
            const S2 b = B::mul(R.y, R.y), j = B::mul(R.x, R.x), xy = B::mul(R.x, R.y), h = B::mul(R.y, R.z, 2);
            const S2 E = B::mul(R.z, R.zt), F = B::mul(R.z, R.zt, 3), ht = B::mul(R.y, R.zt, 2);
            const S2 hx = B::mul(R.y, R.zx, 2), hy = B::mul(R.y, R.zy, 2);
            Line ln;
            for (int p = 0; p < 3; p++) ln.P[p][0] = B::mul(h, Y_p[p], -1);           // −h·(−A.y)·ξ^p
            for (int p = 0; p < 4; p++) ln.P[p][1] = B::mul(j, XA_p[p], 3);            // 3j·A.x·ξ^p
            for (int p = 0; p < 4; p++) ln.P[p][3] = p == 0 ? mat(E - b) : B::mul(E - b, XI_p[p]);
            RPt N;
            { Acc2 a(*this); a.add(xy, b, 2); a.add(xy, F, -2); N.x = a.emit(); }
            { Acc2 a(*this); a.add(b, b); a.add(F, b, 2); a.add(F, E, -1); N.y = a.emit(); }
            N.z = B::mul(b, h, 4);
            N.zt = B::mul(b, ht, 4);
            N.zx = B::mul(b, hx, 4); N.zy = B::mul(b, hy, 4);
            R = N;
            return ln;
        };
        // R ← R + Q (Q affine, with zx = x_Q·z, zy = y_Q·z at hand); xq_p / yq_p: x_Q·ξ^p, y_Q·ξ^p
        auto add_step = [&](const std::vector<S2>& xq_p, const std::vector<S2>& yq_p) {
            const S2 th = R.y - R.zy, la = R.x - R.zx;
            const S2 c = B::mul(th, th), d = B::mul(la, la), lz = B::mul(la, R.z), lx = B::mul(la, R.x), tx = B::mul(th, R.x), tl = B::mul(th, la),
                     tz = B::mul(th, R.z), ly = B::mul(la, R.y), lzt = B::mul(la, R.zt), lzx = B::mul(la, R.zx), lzy = B::mul(la, R.zy);
            Line ln;
            for (int p = 0; p < 3; p++) ln.P[p][0] = B::mul(la, Y_p[p]);               // λ·(−A.y)·ξ^p
            for (int p = 0; p < 4; p++) ln.P[p][1] = B::mul(th, XA_p[p], -1);          // −θ·A.x·ξ^p
            for (int p = 0; p < 4; p++) { Acc2 a(*this); a.add(th, xq_p[p]); a.add(la, yq_p[p], -1); ln.P[p][3] = a.emit(); }
            RPt N;
            { Acc2 a(*this); a.add(d, d); a.add(lz, c); a.add(lx, d, -2); N.x = a.emit(); }
            { Acc2 a(*this); a.add(tx, d, 3); a.add(tl, d, -1); a.add(tz, c, -1); a.add(ly, d, -1); N.y = a.emit(); }
            N.z = B::mul(lz, d);
            N.zt = B::mul(lzt, d);
            N.zx = B::mul(lzx, d); N.zy = B::mul(lzy, d);
            R = N;
            return ln;
        };
        auto absorb = [&](const Line& ln) {
            const S12 L = mul_line(fixed_M(n_fixed++), ln);
            f = mul12(f, L, true);
        };
        split_threshold = getenv("PVM_T1") ? atoi(getenv("PVM_T1")) : 3;   // the f chain is the critical one: its sums are spread over lanes
        for (int i = 63; i >= 0; i--) {
            const bool bit = (ATE_LOW >> i) & 1;
            if (i != 63) f = sqr12(f);
            absorb(dbl_step());
            if (bit) absorb(add_step(XQ_p, YQ_p));
        }
        {   // Frobenius corrections: Q1 = ψ(B), Q2 = −ψ²(B)  (tower.cuh miller_loop_groth16)
            const S2 G2c = cfq2(pt.gamma2), G3c = cfq2(pt.gamma3);
            const S2 q1x = B::mul(conj(xB), G2c), q1y = B::mul(conj(yB), G3c);
            const S2 q2x = B::mul(conj(q1x), G2c), q2y = B::mul(conj(q1y), G3c, -1);
            for (int k = 0; k < 2; k++) {
                const S2 qx = k == 0 ? q1x : q2x, qy = k == 0 ? q1y : q2y;
                R.zx = B::mul(qx, R.z); R.zy = B::mul(qy, R.z);
                absorb(add_step(xi_powers(qx, 4), xi_powers(qy, 4)));
            }
        }
        {   // × the key's own Miller value of (α, β)
            S12 K;
            const Fq12& m = vk.ml_alpha_beta;
            const Fq2 kc[6] = {m.c0.c0, m.c1.c0, m.c0.c1, m.c1.c1, m.c0.c2, m.c1.c2};
            for (int k = 0; k < 6; k++) { K.c[k] = cfq2(kc[k]); K.x[k] = cfq2(kc[k] * xi_); K.x2[k] = cfq2(kc[k] * xi_ * xi_); }
            K.has_x = K.has_x2 = true;
            f = mul12(f, K, true);
        }
        split_threshold = getenv("PVM_T2") ? atoi(getenv("PVM_T2")) : 3;

        // ------------------------------------------------------------------ final exponentiation (tower.cuh final_exponentiation)
        S12 t1;
        {   // f^(q⁶−1) = conj(f)·f⁻¹; f = g + h·w with g, h ∈ Fq6 = Fq2[v]/(v³ − ξ), f⁻¹ = (g − h·w)/(g² − v·h²)
            const S2 g0 = f.c[0], g1 = f.c[2], g2 = f.c[4], h0 = f.c[1], h1 = f.c[3], h2 = f.c[5];
            const S2 xg1 = f.x[2], xg2 = f.x[4], xh1 = f.x[3], xh2 = f.x[5];
            S2 D0, D1, D2;
            { Acc2 a(*this); a.add(g0, g0); a.add(g1, xg2, 2); a.add(h0, xh2, -2); a.add(h1, xh1, -1); D0 = a.emit(); }
            { Acc2 a(*this); a.add(g0, g1, 2); a.add(g2, xg2); a.add(h0, h0, -1); a.add(h1, xh2, -2); D1 = a.emit(); }
            { Acc2 a(*this); a.add(g0, g2, 2); a.add(g1, g1); a.add(h0, h1, -2); a.add(h2, xh2, -1); D2 = a.emit(); }
            const S2 xD1 = mat(mul_xi(D1)), xD2 = mat(mul_xi(D2));
            S2 A, Bv, C, Fv;
            { Acc2 a(*this); a.add(D0, D0); a.add(D1, xD2, -1); A = a.emit(); }
            { Acc2 a(*this); a.add(D2, xD2); a.add(D0, D1, -1); Bv = a.emit(); }
            { Acc2 a(*this); a.add(D1, D1); a.add(D0, D2, -1); C = a.emit(); }
            { Acc2 a(*this); a.add(xD2, Bv); a.add(xD1, C); a.add(D0, A); Fv = a.emit(); }
            const Val nrm = dot({{Fv.a, Fv.a}, {Fv.b, Fv.b}});
            const Val ninv = pow_fixed(nrm, e_inv);
            const S2 Fi = {B::mul(Fv.a, ninv), B::mul(Fv.b * -1, ninv)};
            const S2 E0 = B::mul(A, Fi), E1 = B::mul(Bv, Fi), E2 = B::mul(C, Fi);   // (g² − v·h²)⁻¹
            S12 fi;
            { Acc2 a(*this); a.add(g0, E0); a.add(xg1, E2); a.add(xg2, E1); fi.c[0] = a.emit(); }
            { Acc2 a(*this); a.add(g0, E1); a.add(g1, E0); a.add(xg2, E2); fi.c[2] = a.emit(); }
            { Acc2 a(*this); a.add(g0, E2); a.add(g1, E1); a.add(g2, E0); fi.c[4] = a.emit(); }
            { Acc2 a(*this); a.add(h0, E0, -1); a.add(xh1, E2, -1); a.add(xh2, E1, -1); fi.c[1] = a.emit(); }
            { Acc2 a(*this); a.add(h0, E1, -1); a.add(h1, E0, -1); a.add(xh2, E2, -1); fi.c[3] = a.emit(); }
            { Acc2 a(*this); a.add(h0, E2, -1); a.add(h1, E1, -1); a.add(h2, E0, -1); fi.c[5] = a.emit(); }
            t1 = mulx(fi, conj12(f), true);
        }
        t1 = mulx(frob2(t1, true), t1, true);                 // ^(q²+1)
        ensure_x2(t1);
        const S12 fp = frob1(t1, true), fp2 = frob2(t1, true), fp3 = frob1(fp2, true);
        const S12 fu = pow_u(t1), fu2 = pow_u(fu), fu3 = pow_u(fu2);
        S12 y3 = frob1(fu, true);
        const S12 fu2p = frob1(fu2, true), fu3p = frob1(fu3, true), y2 = frob2(fu2, true);
        const S12 y0 = mulx(mulx(fp, fp2, true), fp3, true);
        const S12 y1 = conj12(t1), y5 = conj12(fu2);
        y3 = conj12(y3);
        const S12 y4 = conj12(mulx(fu, fu2p, true));
        const S12 y6 = conj12(mulx(fu3, fu3p, true));
        S12 t0 = mulx(mulx(cyc_sqr(y6), y4, true), y5, true);
        S12 t2 = mulx(mulx(y3, y5, true), t0, true);
        t0 = mulx(t0, y2, true);
        t2 = cyc_sqr(mulx(cyc_sqr(t2), t0, true));
        t0 = mulx(t2, y1, true);
        t2 = mulx(t2, y0, true);
        const S12 res = mulx(cyc_sqr(t0), t2, false);

        // ------------------------------------------------------------------ G2 membership of B (tower.cuh g2_in_subgroup)
        struct XZ { S2 X, Y, ZZ, ZZZ; };
        auto g2_dbl = [&](const XZ& p) {
            const S2 V = B::mul(p.Y, p.Y, 4), M = B::mul(p.X, p.X, 3);
            const S2 W = B::mul(p.Y, V, 2), S = B::mul(p.X, V), MM = B::mul(M, M);
            XZ r;
            r.ZZ = B::mul(V, p.ZZ);
            r.X = mat(MM - S * 2);
            { Acc2 a(*this); a.add(M, S, 3); a.add(M, MM, -1); a.add(W, p.Y, -1); r.Y = a.emit(); }
            r.ZZZ = B::mul(W, p.ZZZ);
            return r;
        };
        // p + q with q = (X2, Y2, ZZ2, ZZZ2); affine q: ZZ2 = ZZZ2 = 1
        auto g2_add = [&](const XZ& p, const XZ& q, bool q_affine) {
            const S2 U1 = q_affine ? p.X : B::mul(p.X, q.ZZ), S1 = q_affine ? p.Y : B::mul(p.Y, q.ZZZ);
            const S2 U2 = B::mul(q.X, p.ZZ), S2v = B::mul(q.Y, p.ZZZ);
            const S2 zz12 = q_affine ? p.ZZ : B::mul(p.ZZ, q.ZZ), zzz12 = q_affine ? p.ZZZ : B::mul(p.ZZZ, q.ZZZ);
            const S2 Pl = U2 - U1, Rl = S2v - S1;
            const S2 PP = B::mul(Pl, Pl), RR = B::mul(Rl, Rl), Pd = mat(Pl), Rr = mat(Rl);
            const S2 PPP = B::mul(Pd, PP), Q = B::mul(U1, PP);
            XZ r;
            r.ZZ = B::mul(zz12, PP);
            r.X = mat(RR - PPP - Q * 2);
            { Acc2 a(*this); a.add(Rr, Q, 3); a.add(Rr, RR, -1); a.add(Rr, PPP); a.add(S1, PPP, -1); r.Y = a.emit(); }
            r.ZZZ = B::mul(zzz12, PPP);
            return r;
        };
        const u64 X = 4965661367192848881ULL;
        const XZ Paff = {xB, yB, one2(), one2()};
        XZ xP = Paff;
        for (int i = 61; i >= 0; i--) {
            xP = g2_dbl(xP);
            if ((X >> i) & 1) xP = g2_add(xP, Paff, true);
        }
        // ψ^k on XYZZ coordinates: (X, Y, ZZ, ZZZ) ↦ (conj^k(X)·A_k, conj^k(Y)·B_k, conj^k(ZZ), conj^k(ZZZ))
        Fq2 Ak[4], Bk[4];
        Ak[0] = Fq2::one(); Bk[0] = Fq2::one();
        for (int k = 1; k < 4; k++) { Ak[k] = Ak[k - 1].conj() * pt.gamma2; Bk[k] = Bk[k - 1].conj() * pt.gamma3; }
        auto psi_k = [&](const XZ& p, int k) {
            XZ r;
            auto cj = [&](const S2& s) { return (k & 1) ? conj(s) : s; };
            r.X = B::mul(cj(p.X), cfq2(Ak[k])); r.Y = B::mul(cj(p.Y), cfq2(Bk[k]));
            r.ZZ = mat(cj(p.ZZ)); r.ZZZ = mat(cj(p.ZZZ));
            return r;
        };
        XZ lhs = g2_add(xP, Paff, true);                     // [x+1]P
        lhs = g2_add(lhs, psi_k(xP, 1), false);
        lhs = g2_add(lhs, psi_k(xP, 2), false);
        const XZ rhs = psi_k(g2_dbl(xP), 3);

        // ------------------------------------------------------------------ the verdict
        std::vector<Val> fin;
        for (int k = 0; k < 6; k++) { fin.push_back(res.c[k].a); fin.push_back(res.c[k].b); }
        const S2 k0 = B::mul(lhs.X, rhs.ZZ), k1 = B::mul(rhs.X, lhs.ZZ), k2 = B::mul(lhs.Y, rhs.ZZZ), k3 = B::mul(rhs.Y, lhs.ZZZ);
        for (const S2* s : {&k0, &k1, &k2, &k3}) { fin.push_back(s->a); fin.push_back(s->b); }
        fin.push_back(lhs.ZZ.a); fin.push_back(lhs.ZZ.b); fin.push_back(rhs.ZZ.a); fin.push_back(rhs.ZZ.b);
        if ((int)fin.size() != FA_COUNT) throw std::runtime_error("pvm: SP_FINAL layout");
        special(SP_FINAL, fin, 0);
        Program P = schedule();
        exp_words(P.exps[EXP_SQRT], 1); exp_words(P.exps[EXP_INV_SQRT], -3); exp_words(P.exps[EXP_INV], -2);
        return P;
    }
};

inline Program build_verify_program(const PairingTables& pt, const VerifyKeyHost& vk) {
    ProgramBuilder b(pt);
    return b.build(vk);
}

}  // namespace pvm
}  // namespace zk
