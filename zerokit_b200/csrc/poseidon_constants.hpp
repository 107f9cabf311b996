// Host-side generation of the Poseidon round constants and MDS matrices from the Grain LFSR, as the
// reference does at start-up (utils/src/poseidon/poseidon_constants.rs:15-263 via
// rln/src/hashers.rs:14-26).  One-off initialisation work; the tables are then uploaded to GPU
// constant memory.  Bit generation: 80-bit state, taps {62,51,38,23,13,0}, 160 warm-up clocks,
// output bits taken in pairs (keep the 2nd bit when the 1st is 1).
#pragma once
#include <vector>

#include "poseidon.cuh"

namespace zk {

class GrainStream {
   public:
    GrainStream(u32 field_bits, u32 t, u32 rf, u32 rp) {
        for (bool& b : s_) b = false;
        s_[1] = true;  // prime field; S-box x^α (bits 2..5 stay 0)
        put(6, 17, field_bits);
        put(18, 29, t);
        put(30, 39, rf);
        put(40, 49, rp);
        for (int i = 50; i < 80; i++) s_[i] = true;
        for (int i = 0; i < 160; i++) clock();
    }
    // next 254-bit integer, first emitted bit most significant; out = 8 little-endian words
    void next254(u32* out) {
        for (int i = 0; i < 8; i++) out[i] = 0;
        for (int n = 0; n < 254; n++) {
            bool first = clock();
            while (!first) {
                clock();
                first = clock();
            }
            u32 bit = clock() ? 1u : 0u;
            for (int i = 7; i > 0; i--) out[i] = (out[i] << 1) | (out[i - 1] >> 31);
            out[0] = (out[0] << 1) | bit;
        }
    }

   private:
    void put(int lo, int hi, u32 v) {
        for (int i = hi; i >= lo; i--) {
            s_[i] = v & 1;
            v >>= 1;
        }
    }
    bool clock() {
        bool nb = s_[(h_ + 62) % 80] ^ s_[(h_ + 51) % 80] ^ s_[(h_ + 38) % 80] ^ s_[(h_ + 23) % 80] ^ s_[(h_ + 13) % 80] ^ s_[h_];
        s_[h_] = nb;
        h_ = (h_ + 1) % 80;
        return nb;
    }
    bool s_[80];
    int h_ = 0;
};

// ark: (rf+rp)*t elements by rejection sampling; mds[i][j] = 1/(x_i + y_j) with x, y reduced mod r
inline void poseidon_generate(u32 t, u32 rf, u32 rp, Fr* ark, Fr* mds) {
    GrainStream g(254, t, rf, rp);
    u32 v[8], p[8];
    for (int i = 0; i < 8; i++) p[i] = FrCfg::p(i);
    u32 n = 0;
    while (n < (rf + rp) * t) {
        g.next254(v);
        if (Fr::raw_cmp(v, p) < 0) ark[n++] = Fr::from_canonical(v);
    }
    std::vector<Fr> xs(t), ys(t);
    auto mod_p = [&]() {
        g.next254(v);
        while (Fr::raw_cmp(v, p) >= 0) Fr::raw_sub(v, v, p);
        return Fr::from_canonical(v);
    };
    for (auto& x : xs) x = mod_p();
    for (auto& y : ys) y = mod_p();
    for (u32 i = 0; i < t; i++)
        for (u32 j = 0; j < t; j++) mds[i * t + j] = (xs[i] + ys[j]).inv();
}

inline void poseidon_fill_tables(PoseidonTables& pt) {
    poseidon_generate(2, 8, 56, pt.ark2, pt.mds2);
    poseidon_generate(3, 8, 57, pt.ark3, pt.mds3);
    poseidon_generate(4, 8, 56, pt.ark4, pt.mds4);
}

}  // namespace zk
