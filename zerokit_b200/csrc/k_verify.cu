// Groth16 verification on the device, one proof per thread.
//
// Replaces, for the drop-in boundary, rln/src/protocol/proof.rs:856-894 (verify_zk_proof:
// public inputs [y, root, nullifier, x, external_nullifier] → prepare_verifying_key +
// Groth16::verify_proof, ark-groth16 0.5.0) and ark-serialize's Proof::deserialize_compressed
// (rln/src/protocol/proof.rs:456-470).  Check performed:
//     e(−A, B) · e(α, β) · e(vk_x, γ) · e(C, δ) == 1,   vk_x = γ_abc[0] + Σ xᵢ·γ_abc[i+1]
// with a single shared final exponentiation.
#include "device_api.hpp"
#include "fixed_base.cuh"

namespace zk {

__constant__ PairingTables c_pair;

void pairing_upload_tables(const PairingTables& t) { ZK_CUDA_CHECK(cudaMemcpyToSymbol(c_pair, &t, sizeof(PairingTables))); }

__device__ __forceinline__ void load_words(const uint8_t* p, u32* w) {
    // proof bytes are only 1-byte aligned inside the wire format
    for (int i = 0; i < 8; i++) w[i] = (u32)p[4 * i] | ((u32)p[4 * i + 1] << 8) | ((u32)p[4 * i + 2] << 16) | ((u32)p[4 * i + 3] << 24);
}
__device__ __forceinline__ bool fq_canonical_ok(const u32* w) {
    u32 q[8];
    for (int i = 0; i < 8; i++) q[i] = FqCfg::p(i);
    return Fq::raw_cmp(w, q) < 0;
}
__device__ __forceinline__ bool fq_larger_half(const Fq& y) {
    u32 c[8], q[8], n[8];
    y.to_canonical(c);
    for (int i = 0; i < 8; i++) q[i] = FqCfg::p(i);
    Fq::raw_sub(n, q, c);
    return Fq::raw_cmp(c, n) > 0;
}
// square root in Fq (q ≡ 3 mod 4): a^((q+1)/4); returns false if a is a non-residue
__device__ bool fq_sqrt(const Fq& a, Fq& out) {
    u32 e[8];
    for (int i = 0; i < 8; i++) e[i] = FqCfg::p(i);
    e[0] += 1;  // low word of q is 0xd87cfd47: no carry
    for (int i = 0; i < 7; i++) e[i] = (e[i] >> 2) | (e[i + 1] << 30);
    e[7] >>= 2;
    Fq r = a.pow(e);
    out = r;
    return r.sqr() == a;
}
// square root in Fq2 (complex method); returns false if none exists
__device__ bool fq2_sqrt(const Fq2& a, Fq2& out) {
    if (a.is_zero()) { out = a; return true; }
    if (a.b.is_zero()) {
        Fq r;
        if (fq_sqrt(a.a, r)) { out = {r, Fq::zero()}; return true; }
        if (fq_sqrt(a.a.neg(), r)) { out = {Fq::zero(), r}; return true; }  // (r·u)² = −r²
        return false;
    }
    Fq alpha;
    if (!fq_sqrt(a.a.sqr() + a.b.sqr(), alpha)) return false;
    Fq delta = halve(a.a + alpha);
    Fq x0;
    if (!fq_sqrt(delta, x0)) {
        delta = halve(a.a - alpha);
        if (!fq_sqrt(delta, x0)) return false;
    }
    Fq x1 = a.b * (x0.dbl()).inv();
    out = {x0, x1};
    return out.sqr() == a;
}

// returns 0 ok, 1 malformed
__device__ int decompress_g1(const uint8_t* b, G1Affine& p) {
    u32 x[8];
    load_words(b, x);
    const u32 flags = x[7] >> 30;
    x[7] &= 0x3fffffffu;
    if (flags == 3) return 1;
    if (!fq_canonical_ok(x)) return 1;
    if (flags & 1) {  // infinity: ark-serialize reads x as a canonical field element and then ignores it (ark-ec short_weierstrass
        p = G1Affine::infinity();   // deserialize_with_mode: `if flags.is_infinity() { Self::identity() }`)
        return 0;
    }
    Fq X = Fq::from_canonical(x), y;
    if (!fq_sqrt(X.sqr() * X + Fq::from_u32(3), y)) return 1;
    const bool want_larger = (flags & 2) != 0;
    if (fq_larger_half(y) != want_larger) y = y.neg();
    p = {X, y};
    return 0;
}
__device__ int decompress_g2(const uint8_t* b, G2Affine& p) {
    u32 x0[8], x1[8];
    load_words(b, x0);
    load_words(b + 32, x1);
    const u32 flags = x1[7] >> 30;
    x1[7] &= 0x3fffffffu;
    if (flags == 3) return 1;
    if (!fq_canonical_ok(x0) || !fq_canonical_ok(x1)) return 1;
    if (flags & 1) {   // infinity with any canonical x, as ark-serialize accepts it
        p = G2Affine::infinity();
        return 0;
    }
    Fq2 X = {Fq::from_canonical(x0), Fq::from_canonical(x1)};
    Fq2 y;
    if (!fq2_sqrt(X.sqr() * X + c_pair.twist_b, y)) return 1;
    const bool larger = y.b.is_zero() ? fq_larger_half(y.a) : fq_larger_half(y.b);
    if (larger != ((flags & 2) != 0)) y = y.neg();
    p = {X, y};
    if (!g2_in_subgroup(&c_pair, p)) return 1;   // G2 has a cofactor; ark validates on deserialisation
    return 0;
}

__device__ __forceinline__ void store_fq_words(uint8_t* p, const Fq& v) {
    u32 c[8];
    v.to_canonical(c);
    for (int i = 0; i < 8; i++) reinterpret_cast<u32*>(p)[i] = c[i];
}

__global__ void __launch_bounds__(32) k_decompress(const uint8_t* __restrict__ proofs, size_t n, uint8_t* __restrict__ affine,
                                                   uint8_t* __restrict__ ok) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    G1Affine A, C;
    G2Affine Bp;
    int bad = decompress_g1(proofs + 128 * j, A) | decompress_g2(proofs + 128 * j + 32, Bp) | decompress_g1(proofs + 128 * j + 96, C);
    ok[j] = bad ? 0 : 1;
    uint8_t* o = affine + 256 * j;
    for (int i = 0; i < 256; i++) o[i] = 0;
    if (bad) return;
    if (A.is_inf()) o[63] = 0x40; else { store_fq_words(o, A.x); store_fq_words(o + 32, A.y); }
    if (Bp.is_inf()) o[191] = 0x40;
    else { store_fq_words(o + 64, Bp.x.a); store_fq_words(o + 96, Bp.x.b); store_fq_words(o + 128, Bp.y.a); store_fq_words(o + 160, Bp.y.b); }
    if (C.is_inf()) o[255] = 0x40; else { store_fq_words(o + 192, C.x); store_fq_words(o + 224, C.y); }
}

__global__ void __launch_bounds__(32) k_verify(VerifyKeyDev vk, const uint8_t* __restrict__ proofs, const uint8_t* __restrict__ publics,
                                               size_t n, uint8_t* __restrict__ ok) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    G1Affine A, C;
    G2Affine Bp;
    if (decompress_g1(proofs + 128 * j, A) | decompress_g2(proofs + 128 * j + 32, Bp) | decompress_g1(proofs + 128 * j + 96, C)) {
        ok[j] = 2;
        return;
    }
    G1XYZZ vkx = G1XYZZ::from_affine(vk.gamma_abc[0]);
    const size_t per_base = (size_t)vk.gK << (vk.gc - 1);
    for (u32 i = 0; i < vk.n_public; i++) {
        u32 x[8], m[8];
        load_words(publics + (j * vk.n_public + i) * 32, x);
        for (int k = 0; k < 8; k++) m[k] = FrCfg::p(k);
        while (Fr::raw_cmp(x, m) >= 0) Fr::raw_sub(x, x, m);
        vkx.add(fixed_base_mul<Fq>(vk.gamma_tab + i * per_base, vk.gc, vk.gK, x));   // 32 additions instead of 254 doublings
    }
    // one Miller loop for the three proof-dependent pairings (shared squarings, inversion-free steps for the variable point B,
    // precomputed lines for γ₂ and δ₂, sparse products), times the key's own Miller value of (α₁, β₂)
    Fq12 f = miller_loop_groth16(&c_pair, Bp, A.neg(), vk.gamma_lam, vk.gamma_c, vkx.to_affine(), vk.delta_lam, vk.delta_c, C);
    f = f * (*vk.ml_alpha_beta);
    ok[j] = final_exponentiation(&c_pair, f) == Fq12::one() ? 1 : 0;
}

void launch_verify(const VerifyKeyDev& vk, const uint8_t* d_proofs, const uint8_t* d_publics, size_t n, uint8_t* d_ok, cudaStream_t s) {
    if (!n) return;
    k_verify<<<(unsigned)((n + 31) / 32), 32, 0, s>>>(vk, d_proofs, d_publics, n, d_ok);
}
void launch_decompress(const uint8_t* d_proofs, size_t n, uint8_t* d_affine, uint8_t* d_ok, cudaStream_t s) {
    if (!n) return;
    k_decompress<<<(unsigned)((n + 31) / 32), 32, 0, s>>>(d_proofs, n, d_affine, d_ok);
}

}  // namespace zk
