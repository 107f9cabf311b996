// BN254 G1 / G2 group arithmetic for the prover kernels: affine points in HBM (64 B / 128 B, limbs
// in Montgomery form), extended-Jacobian "XYZZ" accumulators in registers (x = X/ZZ, y = Y/ZZZ,
// ZZ³ = ZZZ²) — mixed addition costs 8M+2S and needs no inversion, doubling 6M+4S... (dbl-2008-s-1).
//
// Replaces ark-ec 0.5.0's short-Weierstrass Projective/Affine types used by the reference prover
// (rln/src/partial_proof.rs:98-104,226-273; Cargo.lock:106).  Results are unique group elements,
// so after to_affine() they are bit-identical to ark's.
#pragma once
#include "fp.cuh"

namespace zk {

// ------------------------------------------------------------------------------- Fq2 = Fq[u]/(u²+1)
struct Fq2 {
    Fq a, b;
    static HD Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static HD Fq2 one() { return {Fq::one(), Fq::zero()}; }
    HD bool is_zero() const { return a.is_zero() && b.is_zero(); }
    HD bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
    HD bool operator!=(const Fq2& o) const { return !(*this == o); }
    HD Fq2 operator+(const Fq2& o) const { return {a + o.a, b + o.b}; }
    HD Fq2 operator-(const Fq2& o) const { return {a - o.a, b - o.b}; }
    HD Fq2 neg() const { return {a.neg(), b.neg()}; }
    HD Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
    HD Fq2 conj() const { return {a, b.neg()}; }
    HD Fq2 operator*(const Fq2& o) const {
        // two 2-term dot products, each with a single Montgomery reduction: 4·64 + 2·64 wide MADs — the same count as
        // Karatsuba's three full products, without its five 256-bit additions (and with two reductions instead of three)
#if defined(__CUDA_ARCH__) && defined(ZK_FQ2_KARATSUBA)
        Fq t0 = a * o.a, t1 = b * o.b;
        Fq t2 = (a + b) * (o.a + o.b);
        return {t0 - t1, t2 - t0 - t1};
#else
        return {Fq::dot2(a, o.a, b.neg_lazy(), o.b), Fq::dot2(a, o.b, b, o.a)};
#endif
    }
    // x·y − z·w: two 4-term dot products
    static HD Fq2 sub_prod(const Fq2& x, const Fq2& y, const Fq2& z, const Fq2& w) {
        const Fq nxb = x.b.neg_lazy(), nza = z.a.neg_lazy(), nzb = z.b.neg_lazy();
        const Fq re_a[4] = {x.a, nxb, nza, z.b}, re_b[4] = {y.a, y.b, w.a, w.b};
        const Fq im_a[4] = {x.a, x.b, nza, nzb}, im_b[4] = {y.b, y.a, w.b, w.a};
        return {Fq::dot<4>(re_a, re_b), Fq::dot<4>(im_a, im_b)};
    }
    HD Fq2 sqr() const {
#if defined(__CUDA_ARCH__) && defined(ZK_FQ2_OUTLINE)
        Fq t = Fq::mul_ni(a, b);
        return {Fq::mul_ni(a + b, a - b), t.dbl()};
#else
        Fq t = a * b;
        return {(a + b) * (a - b), t.dbl()};
#endif
    }
    HD Fq2 scale(const Fq& k) const { return {a * k, b * k}; }
    HD Fq2 mul_xi() const {  // × (9 + u)
        Fq a2 = a.dbl(), a4 = a2.dbl(), a8 = a4.dbl();
        Fq b2 = b.dbl(), b4 = b2.dbl(), b8 = b4.dbl();
        return {a8 + a - b, b8 + b + a};
    }
    HD Fq2 inv() const {
        Fq d = (a.sqr() + b.sqr()).inv();
        return {a * d, (b * d).neg()};
    }
    HD Fq2& operator+=(const Fq2& o) { return *this = *this + o; }
    HD Fq2& operator-=(const Fq2& o) { return *this = *this - o; }
    HD Fq2& operator*=(const Fq2& o) { return *this = *this * o; }
};

// Affine point as stored in HBM.  Infinity is encoded as (0, 0) — not on either curve (b ≠ 0).
template <class F>
struct Affine {
    F x, y;
    HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static HD Affine infinity() { return {F::zero(), F::zero()}; }
    HD Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;  // ZZ == 0 ⇒ infinity
    static HD XYZZ infinity() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
    HD bool is_inf() const { return ZZ.is_zero(); }
    static HD XYZZ from_affine(const Affine<F>& p) {
        if (p.is_inf()) return infinity();
        return {p.x, p.y, F::one(), F::one()};
    }
    HD XYZZ neg() const { return {X, Y.neg(), ZZ, ZZZ}; }

    HDN XYZZ dbl() const {  // dbl-2008-s-1 (a = 0)
        if (is_inf()) return *this;
        F U = Y.dbl();
        if (U.is_zero()) return infinity();
        F V = U.sqr();
        F W = U * V;
        F S = X * V;
        F XX = X.sqr();
        F M = XX.dbl() + XX;
        F X3 = M.sqr() - S.dbl();
        F Y3 = F::sub_prod(M, S - X3, W, Y);
        return {X3, Y3, V * ZZ, W * ZZZ};
    }
    // mixed addition acc += p (affine, must not be infinity); complete: handles acc = ∞, p = ±acc
    HD void add_affine(const Affine<F>& p) {
        if (is_inf()) {
            X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        F U2 = p.x * ZZ;
        F S2 = p.y * ZZZ;
        F P = U2 - X;
        F R = S2 - Y;
        if (P.is_zero()) {
            if (R.is_zero()) *this = from_affine(p).dbl();
            else *this = infinity();
            return;
        }
        F PP = P.sqr();
        F PPP = P * PP;
        F Qv = X * PP;
        F X3 = R.sqr() - PPP - Qv.dbl();
        Y = F::sub_prod(R, Qv - X3, Y, PPP);
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }
    HDN void add(const XYZZ& o) {  // add-2008-s
        if (o.is_inf()) return;
        if (is_inf()) { *this = o; return; }
        F U1 = X * o.ZZ, U2 = o.X * ZZ;
        F S1 = Y * o.ZZZ, S2 = o.Y * ZZZ;
        F P = U2 - U1, R = S2 - S1;
        if (P.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = infinity();
            return;
        }
        F PP = P.sqr(), PPP = P * PP, Qv = U1 * PP;
        F X3 = R.sqr() - PPP - Qv.dbl();
        Y = F::sub_prod(R, Qv - X3, S1, PPP);
        X = X3;
        ZZ = ZZ * o.ZZ * PP;
        ZZZ = ZZZ * o.ZZZ * PPP;
    }
    // scalar given as canonical little-endian 8x32 integer
    HDN XYZZ mul(const u32* k) const {
        XYZZ r = infinity();
        for (int i = 255; i >= 0; i--) {
            r = r.dbl();
            if ((k[i >> 5] >> (i & 31)) & 1) r.add(*this);
        }
        return r;
    }
    HDN Affine<F> to_affine() const {
        if (is_inf()) return Affine<F>::infinity();
        F zi = ZZZ.inv();            // 1/ZZZ
        F zz_inv = (zi * ZZ).sqr();  // (ZZ/ZZZ)² = 1/Z² = 1/ZZ
        return {X * zz_inv, Y * zi};
    }
};

typedef Affine<Fq> G1Affine;
typedef Affine<Fq2> G2Affine;
typedef XYZZ<Fq> G1XYZZ;
typedef XYZZ<Fq2> G2XYZZ;

}  // namespace zk
