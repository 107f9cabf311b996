// Device helpers shared by the fixed-base kernels (k_msm_fixed.cu) and the verifier (k_verify.cu): table-entry loads, signed
// window digits, and k·P for a point with a precomputed window table.
#pragma once
#include "curve.cuh"

namespace zk {

template <class F>
__device__ __forceinline__ Affine<F> ld_point(const Affine<F>* p);
template <>
__device__ __forceinline__ Affine<Fq> ld_point<Fq>(const Affine<Fq>* p) {
    return {ldg_fp(&p->x), ldg_fp(&p->y)};
}
template <>
__device__ __forceinline__ Affine<Fq2> ld_point<Fq2>(const Affine<Fq2>* p) {
    return {{ldg_fp(&p->x.a), ldg_fp(&p->x.b)}, {ldg_fp(&p->y.a), ldg_fp(&p->y.b)}};
}

// signed window digit k of the canonical scalar s (with incoming carry); returns digit in [−2^{c−1}, 2^{c−1}]
__device__ __forceinline__ int window_digit(const u32* s, int k, int c, u32& carry) {
    const int bit = k * c;
    const int w = bit >> 5, sh = bit & 31;
    u32 v = 0;
    if (w < 8) {
        v = s[w] >> sh;
        if (sh + c > 32 && w + 1 < 8) v |= s[w + 1] << (32 - sh);
    }
    int d = (int)(v & ((1u << c) - 1)) + (int)carry;
    if (d > (1 << (c - 1))) { d -= (1 << c); carry = 1; } else carry = 0;
    return d;
}

// k·P for the fixed point whose window table is `tb` ([K][2^(c-1)] multiples): K mixed additions, no doublings
template <class F>
__device__ XYZZ<F> fixed_base_mul(const Affine<F>* __restrict__ tb, int c, int K, const u32* k) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    const u32 half = 1u << (c - 1);
    u32 carry = 0;
    for (int w = 0; w < K; w++) {
        int d = window_digit(k, w, c, carry);
        if (d == 0) continue;
        Affine<F> pt = ld_point<F>(tb + (size_t)w * half + ((d < 0 ? -d : d) - 1));
        if (d < 0) pt.y = pt.y.neg();
        acc.add_affine(pt);
    }
    return acc;
}

}  // namespace zk
