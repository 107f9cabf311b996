// Host-side initialisation of the q-dependent pairing constants (Frobenius coefficients and the
// hard-part exponent of the final exponentiation) used by tower.cuh.  The two exponents below are
// public integers derived from the BN254 moduli q and r: (q−1)/6 and (q⁴−q²+1)/r.
#pragma once
#include "tower.cuh"

namespace zk {

inline void pairing_tables_init(PairingTables& pt) {
    static const u32 EXP_Q1_6[8] = {0x2414d4e1u, 0x34b01759u, 0xe6bda1c2u, 0xee9591c2u, 0xc0403964u, 0xf40d60f3u, 0xd032f006u, 0x0810b7bdu};
    static const u32 HARD[24] = {0xccdf42b1u, 0xe81bb482u, 0xf49c36d4u, 0x5abf5cc4u, 0x1da014fdu, 0xf1154e7eu, 0x87cdbacfu, 0xdcc7b44cu, 0x954bcf8au, 0xaaa441e3u, 0xd5095f23u, 0x6b887d56u, 0xf3fd90c6u, 0x79581e16u, 0xd189227du, 0x3b1b1355u, 0x61876f6bu, 0x4e529a58u, 0xd5b12278u, 0x6c0eb522u, 0x83177fafu, 0x331ec151u, 0x0b0759adu, 0x01baaa71u};
    // t = ξ^((q−1)/6); γ2 = t², γ3 = t³; ξ^((q²−1)/6) = t^(q+1) = t·conj(t) ∈ Fq
    Fq2 xi = {Fq::from_u32(9), Fq::from_u32(1)};
    Fq2 t = Fq2::one(), b = xi;
    for (int i = 0; i < 256; i++) {
        if ((EXP_Q1_6[i >> 5] >> (i & 31)) & 1) t = t * b;
        b = b.sqr();
    }
    pt.gamma2 = t.sqr();
    pt.gamma3 = pt.gamma2 * t;
    pt.twist_b = Fq2{Fq::from_u32(3), Fq::zero()} * Fq2{Fq::from_u32(9), Fq::from_u32(1)}.inv();
    pt.frob1[0] = Fq2::one();
    for (int k = 1; k < 6; k++) pt.frob1[k] = pt.frob1[k - 1] * t;
    Fq2 n = t * t.conj();
    pt.frob2[0] = Fq::one();
    for (int k = 1; k < 6; k++) pt.frob2[k] = pt.frob2[k - 1] * n.a;
    for (int i = 0; i < 24; i++) pt.hard[i] = HARD[i];
}

}  // namespace zk
