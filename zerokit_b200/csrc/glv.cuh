// GLV endomorphism helpers for BN254 (host + device): scalar split k = k₁ + k₂·λ and a Straus double multiplication that
// uses it.  Device callers: k_msm_fixed.cu (every MSM term, proof assembly); host callers: tests/host_emul.
#pragma once
#include "curve.cuh"

namespace zk {

namespace glv {
HD u32 S(int i) { return i == 0 ? 0x94d213e3u : 0x89d32568u; }
HD u32 N1(int i) { return i == 0 ? 0x7d4f1128u : i == 1 ? 0x8211bbebu : i == 2 ? 0xeeb859fcu : 0x6f4d8248u; }
HD u32 N2(int i) { return i == 0 ? 0x1221250bu : i == 1 ? 0x0be4e154u : i == 2 ? 0xeeb859fdu : 0x6f4d8248u; }
HD u32 G1C(int i) { return i == 0 ? 0xc7e0b3d7u : i == 1 ? 0xd91d232eu : 0x2u; }   // ⌊2^256·s/r⌋
HD u32 G2C(int i) { return i == 0 ? 0x391eb18du : i == 1 ? 0x7a7bd9d4u : i == 2 ? 0xa773d2cfu : i == 3 ? 0x4ccef014u : 0x2u; }   // ⌊2^256·N₁/r⌋
// β in Montgomery form
HD Fq beta() {
    Fq b;
    b.l[0] = 0xd782e155u; b.l[1] = 0x71930c11u; b.l[2] = 0xffbe3323u; b.l[3] = 0xa6bb947cu;
    b.l[4] = 0xd4741444u; b.l[5] = 0xaa303344u; b.l[6] = 0x26594943u; b.l[7] = 0x2c3b3f0du;
    return b;
}
HD Fq beta2() {   // β² in Montgomery form
    Fq b;
    b.l[0] = 0x13e80b9cu; b.l[1] = 0x3350c88eu; b.l[2] = 0xdb5e56b9u; b.l[3] = 0x7dce557cu;
    b.l[4] = 0xb615564au; b.l[5] = 0x6001b4b8u; b.l[6] = 0x020217e0u; b.l[7] = 0x2682e617u;
    return b;
}
// low NR words of a × b
template <int NA, int NB, int NR>
HD void mul_words(const u32* a, const u32* b, u32* r) {
#pragma unroll
    for (int i = 0; i < NR; i++) r[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; i++) {
        u64 carry = 0;
#pragma unroll
        for (int j = 0; j < NB; j++) {
            if (i + j < NR) {
                u64 t = (u64)a[i] * b[j] + r[i + j] + carry;
                r[i + j] = (u32)t;
                carry = t >> 32;
            }
        }
        if (i + NB < NR) r[i + NB] = (u32)carry;
    }
}
// |k₁| (half = 0) or |k₂| (half = 1) of the canonical scalar k into out[0..7] (upper words zero); returns the sign
HD bool split(const u32* k, int half, u32* out) {
    u32 t1[11], t2[13];
    const u32 g1c[3] = {G1C(0), G1C(1), G1C(2)}, g2c[5] = {G2C(0), G2C(1), G2C(2), G2C(3), G2C(4)};
    const u32 S[2] = {glv::S(0), glv::S(1)}, N1[4] = {glv::N1(0), glv::N1(1), glv::N1(2), glv::N1(3)},
              N2[4] = {glv::N2(0), glv::N2(1), glv::N2(2), glv::N2(3)};
    mul_words<8, 3, 11>(k, g1c, t1);
    mul_words<8, 5, 13>(k, g2c, t2);
    const u32* c1 = t1 + 8;   // ⌊k·G1/2^256⌋ < 2^66
    const u32* c2 = t2 + 8;   // ⌊k·G2/2^256⌋ < 2^128
    u32 a[5], b[5], r[5];
    if (half == 0) {          // k₁ = k − c1·s − c2·N₂   (mod 2^160, |k₁| < 2^128)
        mul_words<3, 2, 5>(c1, S, a);
        mul_words<5, 4, 5>(c2, N2, b);
        u64 br = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            u64 d = (u64)k[i] - a[i] - br;
            r[i] = (u32)d; br = (d >> 32) & 1;
        }
        br = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            u64 d = (u64)r[i] - b[i] - br;
            r[i] = (u32)d; br = (d >> 32) & 1;
        }
    } else {                  // k₂ = c1·N₁ − c2·s
        mul_words<3, 4, 5>(c1, N1, a);
        mul_words<5, 2, 5>(c2, S, b);
        u64 br = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            u64 d = (u64)a[i] - b[i] - br;
            r[i] = (u32)d; br = (d >> 32) & 1;
        }
    }
    const bool neg = (r[4] >> 31) != 0;
    if (neg) {
        u64 c = 1;
#pragma unroll
        for (int i = 0; i < 5; i++) { c += (u64)(~r[i]); r[i] = (u32)c; c >>= 32; }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = r[i];
#pragma unroll
    for (int i = 4; i < 8; i++) out[i] = 0;
    return neg;
}
}  // namespace glv

// kp·P + kq·Q for two per-proof points (s·g_a + r·g1_b of the proof assembly): both scalars are GLV-split, so four 128-bit
// scalars act on P, φ(P), Q, φ(Q); one shared run of 128 doublings with a 15-entry table of their subset sums (Straus):
// ≈ 3 100 products where two double-and-add ladders take ≈ 8 600.  use_q = false leaves Q out (r = 0).
HDN G1XYZZ glv_double_mul(const G1XYZZ& P, const u32* kp, const G1XYZZ& Q, const u32* kq, bool use_q) {
    u32 k[4][8];
    bool neg[4];
    neg[0] = glv::split(kp, 0, k[0]);
    neg[1] = glv::split(kp, 1, k[1]);
    neg[2] = glv::split(kq, 0, k[2]);
    neg[3] = glv::split(kq, 1, k[3]);
    G1XYZZ tab[16];
    tab[0] = G1XYZZ::infinity();
    const Fq beta = glv::beta();
    for (int i = 0; i < 4; i++) {
        G1XYZZ b = i < 2 ? P : Q;
        if (i >= 2 && !use_q) b = G1XYZZ::infinity();
        if (i & 1) b.X = b.X * beta;           // φ(X, Y, ZZ, ZZZ) = (β·X, Y, ZZ, ZZZ)
        if (neg[i]) b = b.neg();
        tab[1 << i] = b;
    }
    for (int m = 3; m < 16; m++) {
        if ((m & (m - 1)) == 0) continue;      // powers of two are the bases themselves
        G1XYZZ t = tab[m & (m - 1)];
        t.add(tab[m & -m]);
        tab[m] = t;
    }
    G1XYZZ acc = G1XYZZ::infinity();
    for (int bit = 127; bit >= 0; bit--) {
        acc = acc.dbl();
        u32 idx = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) idx |= ((k[i][bit >> 5] >> (bit & 31)) & 1u) << i;
        if (idx) acc.add(tab[idx]);
    }
    return acc;
}


}  // namespace zk
