// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit Montgomery limbs (R = 2^256), one field
// element per thread, stored as two 128-bit words so that every HBM access is a 16-byte vector
// load/store.
//
// Replaces the ark-ff 0.5.0 / ark-bn254 0.5.0 field types the reference uses everywhere
// (rln/src/circuit/mod.rs:121-137 type aliases; Cargo.lock:60,128).  The Montgomery residue is
// bit-identical to ark-ff's 4x64 representation, so values can be compared limb for limb.
//
// Multiplication: operand-scanning Montgomery product on two interleaved carry chains ("even" holds
// the products a[2k]*b[i], "odd" the products a[2k+1]*b[i], offset by one 32-bit word) so that every
// mad.lo.cc/madc.hi.cc pair maps onto one IMAD.WIDE with carry-in/out.  The schedule was validated
// word-for-word against a Python model of the PTX carry flag before it was written down here.
// A portable C++ version of every operation exists for host-side unit tests (tests/host_emul).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#define HDN inline __host__ __device__ __noinline__   // heavy routines: one copy per kernel image
#else
#define HD inline
#define DEV inline
#define HDN inline
#endif

typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDA_ARCH__)
#define ZK_PTX 1
#else
#define ZK_PTX 0
#endif

namespace zk {

#if ZK_PTX
DEV u32 ptx_add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 ptx_addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 ptx_addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 ptx_sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 ptx_subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 ptx_subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
DEV u32 ptx_mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 ptx_madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 ptx_madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
DEV u32 ptx_madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#endif

// ---------------------------------------------------------------------------------------------
// Field configurations.  limb(i) constant-folds after unrolling, so moduli become immediates.
struct FrCfg {  // scalar field r (rln/src/circuit/iden3calc/graph.rs:14-15)
    static HD constexpr u32 p(int i) {
        return i == 0 ? 0xf0000001u : i == 1 ? 0x43e1f593u : i == 2 ? 0x79b97091u : i == 3 ? 0x2833e848u
             : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    static HD constexpr u32 r2(int i) {  // 2^512 mod r
        return i == 0 ? 0xae216da7u : i == 1 ? 0x1bb8e645u : i == 2 ? 0xe35c59e3u : i == 3 ? 0x53fe3ab1u
             : i == 4 ? 0x53bb8085u : i == 5 ? 0x8c49833du : i == 6 ? 0x7f4e44a5u : 0x0216d0b1u;
    }
    static HD constexpr u32 one(int i) {  // 2^256 mod r
        return i == 0 ? 0x4ffffffbu : i == 1 ? 0xac96341cu : i == 2 ? 0x9f60cd29u : i == 3 ? 0x36fc7695u
             : i == 4 ? 0x7879462eu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    static constexpr u32 INV = 0xefffffffu;  // −r^{-1} mod 2^32
};
struct FqCfg {  // base field q
    static HD constexpr u32 p(int i) {
        return i == 0 ? 0xd87cfd47u : i == 1 ? 0x3c208c16u : i == 2 ? 0x6871ca8du : i == 3 ? 0x97816a91u
             : i == 4 ? 0x8181585du : i == 5 ? 0xb85045b6u : i == 6 ? 0xe131a029u : 0x30644e72u;
    }
    static HD constexpr u32 r2(int i) {  // 2^512 mod q
        return i == 0 ? 0x538afa89u : i == 1 ? 0xf32cfc5bu : i == 2 ? 0xd44501fbu : i == 3 ? 0xb5e71911u
             : i == 4 ? 0x0a417ff6u : i == 5 ? 0x47ab1effu : i == 6 ? 0xcab8351fu : 0x06d89f71u;
    }
    static HD constexpr u32 one(int i) {  // 2^256 mod q
        return i == 0 ? 0xc58f0d9du : i == 1 ? 0xd35d438du : i == 2 ? 0xf5c70b3du : i == 3 ? 0x0a78eb28u
             : i == 4 ? 0x7879462cu : i == 5 ? 0x666ea36fu : i == 6 ? 0x9a07df2fu : 0x0e0a77c1u;
    }
    static constexpr u32 INV = 0xe4866389u;  // −q^{-1} mod 2^32
};

template <class C>
struct alignas(16) Fp {
    u32 l[8];

    static HD Fp zero() { Fp r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
    static HD Fp one() { Fp r; for (int i = 0; i < 8; i++) r.l[i] = C::one(i); return r; }
    static HD Fp modulus() { Fp r; for (int i = 0; i < 8; i++) r.l[i] = C::p(i); return r; }
    static HD Fp rsquared() { Fp r; for (int i = 0; i < 8; i++) r.l[i] = C::r2(i); return r; }

    HD bool is_zero() const { u32 t = 0; for (int i = 0; i < 8; i++) t |= l[i]; return t == 0; }
    HD bool operator==(const Fp& o) const { u32 t = 0; for (int i = 0; i < 8; i++) t |= l[i] ^ o.l[i]; return t == 0; }
    HD bool operator!=(const Fp& o) const { return !(*this == o); }

    // raw 256-bit helpers -------------------------------------------------------------------
    static HD u32 raw_add(u32* r, const u32* a, const u32* b) {  // returns carry
#if ZK_PTX
        r[0] = ptx_add_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) r[i] = ptx_addc_cc(a[i], b[i]);
        return ptx_addc(0, 0);
#else
        u64 c = 0;
        for (int i = 0; i < 8; i++) { c += (u64)a[i] + b[i]; r[i] = (u32)c; c >>= 32; }
        return (u32)c;
#endif
    }
    static HD u32 raw_sub(u32* r, const u32* a, const u32* b) {  // returns borrow (0/1)
#if ZK_PTX
        r[0] = ptx_sub_cc(a[0], b[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) r[i] = ptx_subc_cc(a[i], b[i]);
        return ptx_subc(0, 0) & 1;
#else
        u64 br = 0;
        for (int i = 0; i < 8; i++) { u64 d = (u64)a[i] - b[i] - br; r[i] = (u32)d; br = (d >> 32) & 1; }
        return (u32)br;
#endif
    }
    // r = a − p if that does not borrow (a ≥ p), else a
    static HD void cond_sub_p(u32* a) {
        u32 t[8], p[8];
#pragma unroll
        for (int i = 0; i < 8; i++) p[i] = C::p(i);
        u32 borrow = raw_sub(t, a, p);
        if (!borrow) {
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = t[i];
        }
    }
    // integer comparison of canonical (non-Montgomery) values
    static HD int raw_cmp(const u32* a, const u32* b) {
        for (int i = 7; i >= 0; i--) { if (a[i] < b[i]) return -1; if (a[i] > b[i]) return 1; }
        return 0;
    }

    HD Fp operator+(const Fp& o) const {
        Fp r;
        raw_add(r.l, l, o.l);  // p < 2^254 ⇒ no carry out
        cond_sub_p(r.l);
        return r;
    }
    HD Fp operator-(const Fp& o) const {
        Fp r;
        u32 borrow = raw_sub(r.l, l, o.l);
        if (borrow) {
            u32 p[8];
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = C::p(i);
            raw_add(r.l, r.l, p);
        }
        return r;
    }
    HD Fp neg() const { return is_zero() ? *this : (zero() - *this); }
    HD Fp dbl() const { return *this + *this; }

    // Montgomery product ----------------------------------------------------------------------
    HD Fp operator*(const Fp& o) const {
        Fp r;
#if ZK_PTX
        mul_ptx(r.l, l, o.l);
#else
        mul_portable(r.l, l, o.l);
#endif
        return r;
    }
    // dedicated squaring: 36 + 64 wide MADs instead of 64 + 64 (the off-diagonal products are taken once against the
    // doubled operand); the rows keep the shape of mul_ptx so the interleaved reduction is unchanged
    HD Fp sqr() const {
#if ZK_PTX
        Fp r;
        sqr_ptx(r.l, l);
        return r;
#else
        return (*this) * (*this);
#endif
    }
    // Σ a[k]·b[k] over N ≤ 5 terms with ONE Montgomery reduction (64·N + 64 wide MADs instead of 128·N).  The a[k] may
    // equal p (callers pass p − y for differences of products); the result is fully reduced:
    // (N·p² + R·p)/R < 2p for N ≤ 5.  Schedule validated on scratch/ptx_model.py (dot2_model and its N-term form).
    template <int N>
    static HD Fp dot(const Fp* a, const Fp* b) {
        static_assert(N >= 1 && N <= 5, "dot: the single conditional subtraction covers at most 5 terms");
        Fp r;
#if ZK_PTX
        u32 E[8], O[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i == 0) row_first_ptx(E, O, a[0].l, b[0].l[0]);
            else row_shift_ptx(E, O, a[0].l, b[0].l[i]);
#pragma unroll
            for (int k = 1; k < N; k++) row_inplace_ptx(E, O, a[k].l, b[k].l[i]);
            reduce_row_ptx(E, O);
        }
        finish_ptx(r.l, E, O);
#else
        mul_portable(r.l, a[0].l, b[0].l);
        for (int k = 1; k < N; k++) { Fp t; mul_portable(t.l, a[k].l, b[k].l); r = r + t; }
#endif
        return r;
    }
    // the same sum for up to 8 terms whose a operands may be any integers below 2^256 (complements p − x, shifted by up to two
    // bits): with W = Σ a[k]/p the result is below (0.19·W + 1)·p, which the caller keeps below 2^256 and finishes reducing
    // (dot_wide subtracts p once).  Inside a row T < (Σ a[k] + p)·2^32 exceeds the 2^288 of E + O·2^32 as soon as W > 4, so
    // these rows carry a ninth word X (weight 2^288).  Schedule validated on scratch/ptx_model_wide.py.  Used by the pairing VM.
    template <int N>
    static HD Fp dot_wide(const Fp* a, const Fp* b) {
        static_assert(N >= 1 && N <= 8, "dot_wide: at most 8 terms");
        Fp r;
#if ZK_PTX
        u32 E[8], O[8], X = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i == 0) row_first_ptx(E, O, a[0].l, b[0].l[0]);
            else row_shift_w_ptx(E, O, X, a[0].l, b[0].l[i]);
#pragma unroll
            for (int k = 1; k < N; k++) row_inplace_w_ptx(E, O, X, a[k].l, b[k].l[i]);
            u32 pw[8];
#pragma unroll
            for (int j = 0; j < 8; j++) pw[j] = C::p(j);
            row_inplace_w_ptx(E, O, X, pw, E[0] * C::INV);   // the reduction row: T += m·p
        }
        finish_ptx(r.l, E, O);   // X is spent by now: the caller's bound keeps the value below 2^256
#else
        mul_portable(r.l, a[0].l, b[0].l);
        for (int k = 1; k < N; k++) { Fp t; mul_portable(t.l, a[k].l, b[k].l); r = r + t; }
#endif
        return r;
    }
    static HD Fp dot2(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
        const Fp x[2] = {a, c}, y[2] = {b, d};
        return dot<2>(x, y);
    }
    // a·b − c·d with one reduction
    static HD Fp sub_prod(const Fp& a, const Fp& b, const Fp& c, const Fp& d) { return dot2(a, b, c.neg_lazy(), d); }
    // p − x as an integer in [1, p] (not reduced: 0 maps to p); only meant as the c operand of dot2
    HD Fp neg_lazy() const {
        Fp r, p;
        for (int i = 0; i < 8; i++) p.l[i] = C::p(i);
        raw_sub(r.l, p.l, l);
        return r;
    }
    // out-of-line product with operands and result in registers (by value): one copy of the multiplier body per kernel
    // image, for Fq2 arithmetic whose fully inlined form overflows the instruction cache (ncu: stall_no_instruction)
    static HDN Fp mul_ni(Fp a, Fp b) { return a * b; }
    HD Fp& operator+=(const Fp& o) { return *this = *this + o; }
    HD Fp& operator-=(const Fp& o) { return *this = *this - o; }
    HD Fp& operator*=(const Fp& o) { return *this = *this * o; }

    static HD void mul_portable(u32* r, const u32* a, const u32* b) {  // CIOS, 64-bit accumulators
        u32 t[10];
        for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            u64 c = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c += (u64)a[j] * b[i] + t[j]; t[j] = (u32)c; c >>= 32; }
            c += t[8]; t[8] = (u32)c; t[9] = (u32)(c >> 32);
            u32 m = t[0] * C::INV;
            c = (u64)m * C::p(0) + t[0];
            c >>= 32;
#pragma unroll
            for (int j = 1; j < 8; j++) { c += (u64)m * C::p(j) + t[j]; t[j - 1] = (u32)c; c >>= 32; }
            c += t[8]; t[7] = (u32)c; t[8] = t[9] + (u32)(c >> 32);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = t[i];
        cond_sub_p(r);
    }


#if ZK_PTX
    // one interleaved reduction row: T += m·p with m chosen so that the low word cancels (T = E + O·2^32)
    static DEV void reduce_row_ptx(u32* E, u32* O) {
        const u32 m = E[0] * C::INV;
        O[0] = ptx_mad_lo_cc(C::p(1), m, O[0]);
        O[1] = ptx_madc_hi_cc(C::p(1), m, O[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            O[j] = ptx_madc_lo_cc(C::p(j + 1), m, O[j]);
            O[j + 1] = ptx_madc_hi_cc(C::p(j + 1), m, O[j + 1]);
        }
        E[0] = ptx_mad_lo_cc(C::p(0), m, E[0]);
        E[1] = ptx_madc_hi_cc(C::p(0), m, E[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            E[j] = ptx_madc_lo_cc(C::p(j), m, E[j]);
            E[j + 1] = ptx_madc_hi_cc(C::p(j), m, E[j + 1]);
        }
        O[7] = ptx_addc(O[7], 0);
    }
    // T += a·bi in place (same shape as the reduction row)
    static DEV void row_inplace_ptx(u32* E, u32* O, const u32* a, u32 bi) {
        O[0] = ptx_mad_lo_cc(a[1], bi, O[0]);
        O[1] = ptx_madc_hi_cc(a[1], bi, O[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            O[j] = ptx_madc_lo_cc(a[j + 1], bi, O[j]);
            O[j + 1] = ptx_madc_hi_cc(a[j + 1], bi, O[j + 1]);
        }
        E[0] = ptx_mad_lo_cc(a[0], bi, E[0]);
        E[1] = ptx_madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            E[j] = ptx_madc_lo_cc(a[j], bi, E[j]);
            E[j + 1] = ptx_madc_hi_cc(a[j], bi, E[j + 1]);
        }
        O[7] = ptx_addc(O[7], 0);
    }
    // the same with the carries of both chains kept in a ninth word X (weight 2^288)
    static DEV void row_inplace_w_ptx(u32* E, u32* O, u32& X, const u32* a, u32 bi) {
        O[0] = ptx_mad_lo_cc(a[1], bi, O[0]);
        O[1] = ptx_madc_hi_cc(a[1], bi, O[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            O[j] = ptx_madc_lo_cc(a[j + 1], bi, O[j]);
            O[j + 1] = ptx_madc_hi_cc(a[j + 1], bi, O[j + 1]);
        }
        X = ptx_addc(X, 0);
        E[0] = ptx_mad_lo_cc(a[0], bi, E[0]);
        E[1] = ptx_madc_hi_cc(a[0], bi, E[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            E[j] = ptx_madc_lo_cc(a[j], bi, E[j]);
            E[j + 1] = ptx_madc_hi_cc(a[j], bi, E[j + 1]);
        }
        O[7] = ptx_addc_cc(O[7], 0);
        X = ptx_addc(X, 0);
    }
    static DEV void row_shift_w_ptx(u32* E, u32* O, u32& X, const u32* a, u32 bi) {
        u32 nE[8], nO[8];
        nE[0] = ptx_add_cc(O[0], E[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            nO[j] = ptx_madc_lo_cc(a[j + 1], bi, E[j + 2]);
            nO[j + 1] = ptx_madc_hi_cc(a[j + 1], bi, E[j + 3]);
        }
        nO[6] = ptx_madc_lo_cc(a[7], bi, 0);
        nO[7] = ptx_madc_hi_cc(a[7], bi, X);
        u32 nX = ptx_addc(0, 0);
        nE[0] = ptx_mad_lo_cc(a[0], bi, nE[0]);
        nE[1] = ptx_madc_hi_cc(a[0], bi, O[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            nE[j] = ptx_madc_lo_cc(a[j], bi, O[j]);
            nE[j + 1] = ptx_madc_hi_cc(a[j], bi, O[j + 1]);
        }
        nO[7] = ptx_addc_cc(nO[7], 0);
        X = ptx_addc(nX, 0);
#pragma unroll
        for (int j = 0; j < 8; j++) { E[j] = nE[j]; O[j] = nO[j]; }
    }
    // T = (T >> 32) + a·bi  (the shift is free: the old E words feed the new O chain and vice versa)
    static DEV void row_shift_ptx(u32* E, u32* O, const u32* a, u32 bi) {
        u32 nE[8], nO[8];
        nE[0] = ptx_add_cc(O[0], E[1]);
#pragma unroll
        for (int j = 0; j < 6; j += 2) {
            nO[j] = ptx_madc_lo_cc(a[j + 1], bi, E[j + 2]);
            nO[j + 1] = ptx_madc_hi_cc(a[j + 1], bi, E[j + 3]);
        }
        nO[6] = ptx_madc_lo_cc(a[7], bi, 0);
        nO[7] = ptx_madc_hi(a[7], bi, 0);
        nE[0] = ptx_mad_lo_cc(a[0], bi, nE[0]);
        nE[1] = ptx_madc_hi_cc(a[0], bi, O[1]);
#pragma unroll
        for (int j = 2; j < 8; j += 2) {
            nE[j] = ptx_madc_lo_cc(a[j], bi, O[j]);
            nE[j + 1] = ptx_madc_hi_cc(a[j], bi, O[j + 1]);
        }
        nO[7] = ptx_addc(nO[7], 0);
#pragma unroll
        for (int j = 0; j < 8; j++) { E[j] = nE[j]; O[j] = nO[j]; }
    }
    static DEV void row_first_ptx(u32* E, u32* O, const u32* a, u32 bi) {
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
            asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(E[j]), "=r"(E[j + 1]) : "r"(a[j]), "r"(bi));
            asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(O[j]), "=r"(O[j + 1]) : "r"(a[j + 1]), "r"(bi));
        }
    }
    static DEV void finish_ptx(u32* r, const u32* E, const u32* O) {  // result = (E >> 32) + O, then one conditional −p
        r[0] = ptx_add_cc(E[1], O[0]);
#pragma unroll
        for (int k = 1; k < 7; k++) r[k] = ptx_addc_cc(E[k + 1], O[k]);
        r[7] = ptx_addc(O[7], 0);
        cond_sub_p(r);
    }
    // schedule validated on scratch/ptx_model.py (sqr_model): row i multiplies a[i] by (a[i], 2·a[i+1..7]) only
    static DEV void sqr_ptx(u32* r, const u32* a) {
        u32 d[8], s1[8];
        d[0] = 0; s1[0] = 0;
#pragma unroll
        for (int j = 1; j < 8; j++) {
            d[j] = __funnelshift_l(a[j - 1], a[j], 1);
            s1[j] = a[j] << 1;
        }
        u32 E[8], O[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const u32 bi = a[i];
            u32 v[8];   // multiplicand words of this row (entries below i are unused)
#pragma unroll
            for (int j = 0; j < 8; j++) v[j] = j == i ? a[i] : (j == i + 1 ? s1[j] : d[j]);
            if (i == 0) {
                row_first_ptx(E, O, v, bi);
            } else {
                u32 nE[8], nO[8];
                nE[0] = ptx_add_cc(O[0], E[1]);
#pragma unroll
                for (int j = 0; j < 6; j += 2) {
                    if (j + 1 >= i) {
                        nO[j] = ptx_madc_lo_cc(v[j + 1], bi, E[j + 2]);
                        nO[j + 1] = ptx_madc_hi_cc(v[j + 1], bi, E[j + 3]);
                    } else {
                        nO[j] = ptx_addc_cc(E[j + 2], 0);
                        nO[j + 1] = ptx_addc_cc(E[j + 3], 0);
                    }
                }
                nO[6] = ptx_madc_lo_cc(v[7], bi, 0);
                nO[7] = ptx_madc_hi(v[7], bi, 0);
                bool started = false;
#pragma unroll
                for (int j = 0; j < 8; j += 2) {
                    if (j >= i) {
                        nE[j] = started ? ptx_madc_lo_cc(v[j], bi, O[j]) : ptx_mad_lo_cc(v[j], bi, O[j]);
                        started = true;
                        nE[j + 1] = ptx_madc_hi_cc(v[j], bi, O[j + 1]);
                    } else {
                        if (j > 0) nE[j] = O[j];
                        nE[j + 1] = O[j + 1];
                    }
                }
                if (started) nO[7] = ptx_addc(nO[7], 0);
#pragma unroll
                for (int j = 0; j < 8; j++) { E[j] = nE[j]; O[j] = nO[j]; }
            }
            reduce_row_ptx(E, O);
        }
        finish_ptx(r, E, O);
    }
    static DEV void mul_ptx(u32* r, const u32* a, const u32* b) {
        // E: words of weight 2^0, 2^32, …  O: words of weight 2^32, 2^64, …  (sum T = E + O·2^32)
        u32 E[8], O[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i == 0) row_first_ptx(E, O, a, b[0]);
            else row_shift_ptx(E, O, a, b[i]);
            reduce_row_ptx(E, O);
        }
        finish_ptx(r, E, O);
    }
#endif

    // conversions -------------------------------------------------------------------------------
    // canonical little-endian integer (must be < p) → Montgomery
    static HD Fp from_canonical(const u32* c) {
        Fp t;
        for (int i = 0; i < 8; i++) t.l[i] = c[i];
        return t * rsquared();
    }
    static HD Fp from_u32(u32 v) {
        u32 c[8] = {v, 0, 0, 0, 0, 0, 0, 0};
        return from_canonical(c);
    }
    HD void to_canonical(u32* c) const {
        Fp o;
        for (int i = 0; i < 8; i++) o.l[i] = 0;
        o.l[0] = 1;
        Fp t = (*this) * o;
        for (int i = 0; i < 8; i++) c[i] = t.l[i];
    }
    // exponent given as canonical little-endian 8x32 integer
    HDN Fp pow(const u32* e) const {
        Fp r = one(), b = *this;
        for (int i = 0; i < 256; i++) {
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * b;
            b = b.sqr();
        }
        return r;
    }
    HDN Fp inv() const {  // a^(p−2); maps 0 to 0
        u32 e[8];
        for (int i = 0; i < 8; i++) e[i] = C::p(i);
        e[0] -= 2;  // p is odd and its low word is > 2
        Fp r = one();
        for (int i = 255; i >= 0; i--) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * (*this);
        }
        return r;
    }
};

typedef Fp<FrCfg> Fr;
typedef Fp<FqCfg> Fq;

// 128-bit global memory access helpers (two 16-byte vector transactions per element)
#if defined(__CUDACC__)
template <class F>
DEV F ld_fp(const F* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    F r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class F>
DEV F ldg_fp(const F* p) {  // read-only path
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    F r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class F>
DEV void st_fp(F* p, const F& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
#endif

}  // namespace zk
