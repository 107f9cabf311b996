// Poseidon permutation over BN254 Fr (x^5 S-box, 8 full rounds, RP partial rounds, dense t×t MDS)
// for state widths t = 2, 3, 4 — the three instantiations the RLN path uses:
//   t=2  id_commitment = H(secret), nullifier = H(a1)         rln/src/protocol/witness.rs:776,814
//   t=3  tree nodes H(l, r), rate_commitment = H(idc, limit)  rln/src/hashers.rs:49-70
//   t=4  a1 = H(secret, ext_nullifier, message_id)            rln/src/protocol/witness.rs:774-775
// Follows utils/src/poseidon/poseidon_hash.rs:63-135 (ark → sbox → mix, output = state[0]).
// Round constants / MDS are produced by the Grain LFSR restated in poseidon_constants.hpp and
// uploaded (Montgomery form) into the PoseidonTables object below.
#pragma once
#include "fp.cuh"

namespace zk {

struct PoseidonTables {
    // (RF+RP)·t round constants, then t·t MDS entries (row-major), per width
    Fr ark2[64 * 2];
    Fr mds2[4];
    Fr ark3[65 * 3];
    Fr mds3[9];
    Fr ark4[64 * 4];
    Fr mds4[16];
};

template <int T>
struct PoseidonShape;
template <>
struct PoseidonShape<2> { static constexpr int RF = 8, RP = 56; };
template <>
struct PoseidonShape<3> { static constexpr int RF = 8, RP = 57; };
template <>
struct PoseidonShape<4> { static constexpr int RF = 8, RP = 56; };

HD Fr sbox5(const Fr& x) {
    Fr x2 = x.sqr();
    return x2.sqr() * x;
}

// state[0] must be 0 and state[1..T-1] the inputs on entry; returns state[0] after the permutation
template <int T>
HD Fr poseidon_permute(Fr* st, const Fr* __restrict__ ark, const Fr* __restrict__ mds) {
    constexpr int RF = PoseidonShape<T>::RF, RP = PoseidonShape<T>::RP;
#pragma unroll 1
    for (int r = 0; r < RF + RP; r++) {
#pragma unroll
        for (int k = 0; k < T; k++) st[k] += ark[r * T + k];
        const bool full = (r < RF / 2) || (r >= RF / 2 + RP);
        st[0] = sbox5(st[0]);
        if (full) {
#pragma unroll
            for (int k = 1; k < T; k++) st[k] = sbox5(st[k]);
        }
        Fr nx[T];
#pragma unroll
        for (int i = 0; i < T; i++) nx[i] = Fr::dot<T>(st, mds + i * T);   // one Montgomery reduction per MDS row
#pragma unroll
        for (int i = 0; i < T; i++) st[i] = nx[i];
    }
    return st[0];
}

HD Fr poseidon1(const PoseidonTables* pt, const Fr& a) {
    Fr st[2] = {Fr::zero(), a};
    return poseidon_permute<2>(st, pt->ark2, pt->mds2);
}
HD Fr poseidon2(const PoseidonTables* pt, const Fr& a, const Fr& b) {
    Fr st[3] = {Fr::zero(), a, b};
    return poseidon_permute<3>(st, pt->ark3, pt->mds3);
}
HD Fr poseidon3(const PoseidonTables* pt, const Fr& a, const Fr& b, const Fr& c) {
    Fr st[4] = {Fr::zero(), a, b, c};
    return poseidon_permute<4>(st, pt->ark4, pt->mds4);
}

}  // namespace zk
