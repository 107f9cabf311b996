// ORACLE (test infrastructure, not product code).
//
// BN254 arithmetic for the C++ CPU restatement of the reference's proving path: 4x64-bit
// Montgomery fields (unsigned __int128), the Fq2/Fq6/Fq12 tower, short-Weierstrass group law
// in Jacobian coordinates, and the optimal-ate pairing.  The reference gets all of this from
// un-vendored crates (ark-ff / ark-ec / ark-bn254 0.5.0, Cargo.lock:60-234); the algorithms
// here are the textbook ones and are validated against oracle/pyref (which is pinned to the
// reference's golden vectors) by tests/test_oracle_cref.py.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

typedef unsigned __int128 u128;
typedef uint64_t u64;

struct U256 {
    u64 l[4];
    bool operator==(const U256& o) const { return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3]; }
};

static inline int u256_cmp(const U256& a, const U256& b) {
    for (int i = 3; i >= 0; i--) {
        if (a.l[i] < b.l[i]) return -1;
        if (a.l[i] > b.l[i]) return 1;
    }
    return 0;
}
static inline u64 u256_add(U256& r, const U256& a, const U256& b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 u256_sub(U256& r, const U256& a, const U256& b) {
    u64 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a.l[i] - b.l[i] - br;
        r.l[i] = (u64)d;
        br = (u64)(d >> 64) & 1;
    }
    return br;
}
static inline bool u256_is_zero(const U256& a) { return (a.l[0] | a.l[1] | a.l[2] | a.l[3]) == 0; }
static inline int u256_bit(const U256& a, int i) { return (a.l[i >> 6] >> (i & 63)) & 1; }

// Field parameter packs -------------------------------------------------------------------
struct FrParams {
    static constexpr U256 MOD = {{0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
};
struct FqParams {
    static constexpr U256 MOD = {{0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL}};
};

template <class P>
struct Fp {
    U256 v;  // Montgomery form (a·2^256 mod p)

    static u64 inv_neg() {  // −p^{-1} mod 2^64
        static u64 c = [] {
            u64 p0 = P::MOD.l[0], x = 1;
            for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
            return (u64)(0 - x);
        }();
        return c;
    }
    static const Fp& R2() {  // 2^512 mod p (Montgomery form of 2^256)
        static Fp c = [] {
            Fp r;
            r.v = {{1, 0, 0, 0}};
            // double 512 times with modular reduction
            for (int i = 0; i < 512; i++) {
                U256 t;
                u64 carry = u256_add(t, r.v, r.v);
                if (carry || u256_cmp(t, P::MOD) >= 0) u256_sub(t, t, P::MOD);
                r.v = t;
            }
            return r;
        }();
        return c;
    }
    static Fp zero() { Fp r; r.v = {{0, 0, 0, 0}}; return r; }
    static const Fp& one() {
        static Fp c = from_u256({{1, 0, 0, 0}});
        return c;
    }
    static Fp from_u256(const U256& a) {  // a must be < p
        Fp r;
        r.v = a;
        return r * R2();
    }
    static Fp from_u64(u64 a) { return from_u256({{a, 0, 0, 0}}); }
    // little-endian bytes, reduced mod p (used for hash outputs / file constants)
    static Fp from_le_bytes_mod(const uint8_t* b, size_t n) {
        Fp acc = zero(), base = from_u64(256);
        for (size_t i = n; i-- > 0;) acc = acc * base + from_u64(b[i]);
        return acc;
    }
    static Fp from_le32(const uint8_t* b) {  // canonical value expected; reduces if not
        U256 a;
        memcpy(a.l, b, 32);
        while (u256_cmp(a, P::MOD) >= 0) u256_sub(a, a, P::MOD);
        return from_u256(a);
    }
    U256 to_u256() const {
        Fp o;
        o.v = {{1, 0, 0, 0}};
        return ((*this) * o).v;
    }
    void to_le32(uint8_t* b) const {
        U256 a = to_u256();
        memcpy(b, a.l, 32);
    }
    bool is_zero() const { return u256_is_zero(v); }
    bool operator==(const Fp& o) const { return v == o.v; }
    bool operator!=(const Fp& o) const { return !(v == o.v); }

    Fp operator+(const Fp& o) const {
        Fp r;
        u64 c = u256_add(r.v, v, o.v);
        if (c || u256_cmp(r.v, P::MOD) >= 0) u256_sub(r.v, r.v, P::MOD);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        if (u256_sub(r.v, v, o.v)) u256_add(r.v, r.v, P::MOD);
        return r;
    }
    Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    Fp dbl() const { return *this + *this; }

    // Montgomery multiplication, the "no-carry" CIOS form ark-ff 0.5 selects for moduli whose top bit is clear (both BN254
    // fields; ark-ff/src/fields/models/fp/montgomery_backend.rs `mul_assign`): the two carry chains of a row (a·b[i] and k·p)
    // are interleaved and the running total never needs a fifth word.  Fully unrolled, −p⁻¹ mod 2^64 folded to a constant.
    static constexpr u64 inv_neg_const() {
        u64 p0 = P::MOD.l[0], x = 1;
        for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
        return (u64)(0 - x);
    }
    Fp operator*(const Fp& o) const {
        constexpr u64 NINV = inv_neg_const();
        constexpr u64 m0 = P::MOD.l[0], m1 = P::MOD.l[1], m2 = P::MOD.l[2], m3 = P::MOD.l[3];
        const u64 a0 = v.l[0], a1 = v.l[1], a2 = v.l[2], a3 = v.l[3];
        u64 r0 = 0, r1 = 0, r2 = 0, r3 = 0;
#pragma GCC unroll 4
        for (int i = 0; i < 4; i++) {
            const u64 bi = o.v.l[i];
            u128 c1 = (u128)a0 * bi + r0;
            const u64 k = (u64)c1 * NINV;
            u128 c2 = (u128)k * m0 + (u64)c1;
            c1 = (u128)a1 * bi + r1 + (u64)(c1 >> 64);
            c2 = (u128)k * m1 + (u64)c1 + (u64)(c2 >> 64);
            r0 = (u64)c2;
            c1 = (u128)a2 * bi + r2 + (u64)(c1 >> 64);
            c2 = (u128)k * m2 + (u64)c1 + (u64)(c2 >> 64);
            r1 = (u64)c2;
            c1 = (u128)a3 * bi + r3 + (u64)(c1 >> 64);
            c2 = (u128)k * m3 + (u64)c1 + (u64)(c2 >> 64);
            r2 = (u64)c2;
            r3 = (u64)(c1 >> 64) + (u64)(c2 >> 64);
        }
        Fp r;
        r.v = {{r0, r1, r2, r3}};
        if (u256_cmp(r.v, P::MOD) >= 0) u256_sub(r.v, r.v, P::MOD);
        return r;
    }
    Fp sqr() const { return (*this) * (*this); }
    Fp& operator+=(const Fp& o) { return *this = *this + o; }
    Fp& operator-=(const Fp& o) { return *this = *this - o; }
    Fp& operator*=(const Fp& o) { return *this = *this * o; }

    Fp pow(const U256& e) const {
        Fp r = one(), b = *this;
        for (int i = 0; i < 256; i++) {
            if (u256_bit(e, i)) r = r * b;
            b = b.sqr();
        }
        return r;
    }
    Fp inv() const {  // a^(p-2); zero maps to zero
        U256 e, two = {{2, 0, 0, 0}};
        u256_sub(e, P::MOD, two);
        return pow(e);
    }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// ------------------------------------------------------------------------------- Fq2 = Fq[u]/(u²+1)
struct Fq2 {
    Fq a, b;  // a + b·u
    static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
    static Fq2 one() { return {Fq::one(), Fq::zero()}; }
    bool is_zero() const { return a.is_zero() && b.is_zero(); }
    bool operator==(const Fq2& o) const { return a == o.a && b == o.b; }
    bool operator!=(const Fq2& o) const { return !(*this == o); }
    Fq2 operator+(const Fq2& o) const { return {a + o.a, b + o.b}; }
    Fq2 operator-(const Fq2& o) const { return {a - o.a, b - o.b}; }
    Fq2 neg() const { return {a.neg(), b.neg()}; }
    Fq2 dbl() const { return {a.dbl(), b.dbl()}; }
    Fq2 conj() const { return {a, b.neg()}; }
    Fq2 operator*(const Fq2& o) const {
        Fq t0 = a * o.a, t1 = b * o.b;
        Fq t2 = (a + b) * (o.a + o.b);
        return {t0 - t1, t2 - t0 - t1};
    }
    Fq2 sqr() const {
        Fq t = a * b;
        return {(a + b) * (a - b), t.dbl()};
    }
    Fq2 scale(const Fq& k) const { return {a * k, b * k}; }
    Fq2 mul_xi() const {  // × (9 + u)
        Fq a2 = a.dbl(), a4 = a2.dbl(), a8 = a4.dbl();
        Fq b2 = b.dbl(), b4 = b2.dbl(), b8 = b4.dbl();
        return {a8 + a - b, b8 + b + a};
    }
    Fq2 inv() const {
        Fq d = (a.sqr() + b.sqr()).inv();
        return {a * d, (b * d).neg()};
    }
    Fq2 pow(const std::vector<u64>& e) const {
        Fq2 r = one(), x = *this;
        for (size_t i = 0; i < e.size() * 64; i++) {
            if ((e[i >> 6] >> (i & 63)) & 1) r = r * x;
            x = x.sqr();
        }
        return r;
    }
    Fq2& operator+=(const Fq2& o) { return *this = *this + o; }
    Fq2& operator-=(const Fq2& o) { return *this = *this - o; }
    Fq2& operator*=(const Fq2& o) { return *this = *this * o; }
};

// ------------------------------------------------------------------------------- Fq6 = Fq2[v]/(v³ − ξ)
struct Fq6 {
    Fq2 c0, c1, c2;
    static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
    static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
    bool operator==(const Fq6& o) const { return c0 == o.c0 && c1 == o.c1 && c2 == o.c2; }
    Fq6 operator+(const Fq6& o) const { return {c0 + o.c0, c1 + o.c1, c2 + o.c2}; }
    Fq6 operator-(const Fq6& o) const { return {c0 - o.c0, c1 - o.c1, c2 - o.c2}; }
    Fq6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    Fq6 operator*(const Fq6& o) const {
        Fq2 t0 = c0 * o.c0, t1 = c1 * o.c1, t2 = c2 * o.c2;
        Fq2 r0 = ((c1 + c2) * (o.c1 + o.c2) - t1 - t2).mul_xi() + t0;
        Fq2 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1 + t2.mul_xi();
        Fq2 r2 = (c0 + c2) * (o.c0 + o.c2) - t0 - t2 + t1;
        return {r0, r1, r2};
    }
    Fq6 mul_v() const { return {c2.mul_xi(), c0, c1}; }  // × v
    Fq6 inv() const {
        Fq2 A = c0.sqr() - (c1 * c2).mul_xi();
        Fq2 B = c2.sqr().mul_xi() - c0 * c1;
        Fq2 C = c1.sqr() - c0 * c2;
        Fq2 F = (c2 * B + c1 * C).mul_xi() + c0 * A;
        Fq2 Fi = F.inv();
        return {A * Fi, B * Fi, C * Fi};
    }
};

// ------------------------------------------------------------------------------- Fq12 = Fq6[w]/(w² − v)
struct Fq12 {
    Fq6 c0, c1;
    static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
    bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
    Fq12 operator*(const Fq12& o) const {
        Fq6 t0 = c0 * o.c0, t1 = c1 * o.c1;
        Fq6 r1 = (c0 + c1) * (o.c0 + o.c1) - t0 - t1;
        return {t0 + t1.mul_v(), r1};
    }
    Fq12 sqr() const { return (*this) * (*this); }
    Fq12 conj() const { return {c0, c1.neg()}; }
    Fq12 inv() const {
        Fq6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * d, (c1 * d).neg()};
    }
};

// ------------------------------------------------------------------------------- curve (Jacobian)
// y² = x³ + b over field F (Fq for G1, Fq2 for G2).
template <class F>
struct Affine {
    F x, y;
    bool inf;
};
template <class F>
struct Jac {
    F X, Y, Z;  // Z == 0 ⇒ infinity
    static Jac infinity() { return {F::one(), F::one(), F::zero()}; }
    bool is_inf() const { return Z.is_zero(); }
    static Jac from_affine(const Affine<F>& p) {
        if (p.inf) return infinity();
        return {p.x, p.y, F::one()};
    }
    Jac dbl() const {
        if (is_inf() || Y.is_zero()) return infinity();
        F A = X.sqr(), B = Y.sqr(), C = B.sqr();
        F t = (X + B).sqr() - A - C;
        F D = t.dbl();
        F E = A.dbl() + A;
        F Fv = E.sqr();
        F X3 = Fv - D.dbl();
        F C8 = C.dbl().dbl().dbl();
        F Y3 = E * (D - X3) - C8;
        F Z3 = (Y * Z).dbl();
        return {X3, Y3, Z3};
    }
    Jac add_affine(const Affine<F>& q) const {  // mixed addition
        if (q.inf) return *this;
        if (is_inf()) return from_affine(q);
        F Z1Z1 = Z.sqr();
        F U2 = q.x * Z1Z1;
        F S2 = q.y * Z * Z1Z1;
        F H = U2 - X, r = S2 - Y;
        if (H.is_zero()) return r.is_zero() ? dbl() : infinity();
        F HH = H.sqr(), HHH = H * HH, V = X * HH;
        F X3 = r.sqr() - HHH - V.dbl();
        F Y3 = r * (V - X3) - Y * HHH;
        F Z3 = Z * H;
        return {X3, Y3, Z3};
    }
    Jac add(const Jac& o) const {
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        F Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        F U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        F S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        F H = U2 - U1, r = S2 - S1;
        if (H.is_zero()) return r.is_zero() ? dbl() : infinity();
        F HH = H.sqr(), HHH = H * HH, V = U1 * HH;
        F X3 = r.sqr() - HHH - V.dbl();
        F Y3 = r * (V - X3) - S1 * HHH;
        F Z3 = Z * o.Z * H;
        return {X3, Y3, Z3};
    }
    Jac neg() const { return {X, Y.neg(), Z}; }
    Jac mul(const U256& k) const {
        Jac r = infinity();
        for (int i = 255; i >= 0; i--) {
            r = r.dbl();
            if (u256_bit(k, i)) r = r.add(*this);
        }
        return r;
    }
    Affine<F> to_affine() const {
        if (is_inf()) return {F::zero(), F::zero(), true};
        F zi = Z.inv(), zi2 = zi.sqr();
        return {X * zi2, Y * zi2 * zi, false};
    }
};
typedef Affine<Fq> G1A;
typedef Affine<Fq2> G2A;
typedef Jac<Fq> G1J;
typedef Jac<Fq2> G2J;

// ------------------------------------------------------------------------------- pairing
// Optimal ate; the running point stays affine on the twist E'(Fq2), lines are embedded into
// Fq12 with u = w⁶ − 9:  l = −yP + (λ·xP)·w + (yR − λ·xR)·w³   (w³ = v·w in the tower).
struct PairingConsts {
    Fq2 gamma2, gamma3;  // ξ^((q−1)/3), ξ^((q−1)/2)
    Fq frob2[6];         // ξ^(k(q²−1)/6), k = 0..5  (all in Fq)
    std::vector<u64> hard;  // (q⁴ − q² + 1)/r
};
const PairingConsts& pairing_consts();  // defined in rln_oracle.cpp

static inline Fq12 line_eval(const Fq2& lam, const G2A& R, const G1A& P) {
    Fq12 l;
    l.c0 = {Fq2{P.y.neg(), Fq::zero()}, Fq2::zero(), Fq2::zero()};
    l.c1 = {lam.scale(P.x), R.y - lam * R.x, Fq2::zero()};
    return l;
}

static inline Fq12 miller_loop(const G2A& Qp, const G1A& P) {
    if (Qp.inf || P.inf) return Fq12::one();
    const PairingConsts& pc = pairing_consts();
    const u64 ATE = 0x9d797039be763ba8ULL;  // 29793968203157093288
    G2A R = Qp;
    Fq12 f = Fq12::one();
    auto dbl_step = [&]() {
        Fq2 xx = R.x.sqr();
        Fq2 lam = (xx.dbl() + xx) * R.y.dbl().inv();
        f = f.sqr() * line_eval(lam, R, P);
        Fq2 x3 = lam.sqr() - R.x.dbl();
        Fq2 y3 = lam * (R.x - x3) - R.y;
        R.x = x3;
        R.y = y3;
    };
    auto add_step = [&](const G2A& S) {
        Fq2 lam = (S.y - R.y) * (S.x - R.x).inv();
        f = f * line_eval(lam, R, P);
        Fq2 x3 = lam.sqr() - R.x - S.x;
        Fq2 y3 = lam * (R.x - x3) - R.y;
        R.x = x3;
        R.y = y3;
    };
    // the python restatement runs i = 63 … 0 starting from f = 1, R = Q (bit 64 of ATE is set
    // and is consumed by the initial R = Q): ATE has 65 bits, top bit index 64.
    for (int i = 63; i >= 0; i--) {
        dbl_step();
        if ((ATE >> i) & 1) add_step(Qp);
    }
    G2A Q1 = {Qp.x.conj() * pc.gamma2, Qp.y.conj() * pc.gamma3, false};
    G2A Q2 = {Q1.x.conj() * pc.gamma2, (Q1.y.conj() * pc.gamma3).neg(), false};
    add_step(Q1);
    add_step(Q2);
    return f;
}

static inline Fq12 frobenius2(const Fq12& f) {
    const PairingConsts& pc = pairing_consts();
    Fq12 r;
    // basis order: c0 = (w⁰, w², w⁴), c1 = (w¹, w³, w⁵)
    r.c0 = {f.c0.c0, f.c0.c1.scale(pc.frob2[2]), f.c0.c2.scale(pc.frob2[4])};
    r.c1 = {f.c1.c0.scale(pc.frob2[1]), f.c1.c1.scale(pc.frob2[3]), f.c1.c2.scale(pc.frob2[5])};
    return r;
}

static inline Fq12 final_exponentiation(const Fq12& f) {
    const PairingConsts& pc = pairing_consts();
    Fq12 t = f.conj() * f.inv();       // f^(q⁶−1)
    t = frobenius2(t) * t;             // ^(q²+1)
    Fq12 r = Fq12::one();              // ^((q⁴−q²+1)/r)
    for (size_t i = pc.hard.size() * 64; i-- > 0;) {
        r = r.sqr();
        if ((pc.hard[i >> 6] >> (i & 63)) & 1) r = r * t;
    }
    return r;
}
