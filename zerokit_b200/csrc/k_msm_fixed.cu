// Batched fixed-base multi-scalar multiplication and Groth16 proof assembly (sm_100a).
//
// Path covered: the five MSMs of ark-groth16's create_proof_with_assignment
// as restated in rln/src/partial_proof.rs:226-273 (A, B in G1, B in G2, L, H) and the final
//   g_a  = α₁ + Σ wᵢ·Aᵢ + r·δ₁
//   g1_b = β₁ + Σ wᵢ·B¹ᵢ + s·δ₁          (only when r ≠ 0, partial_proof.rs:242-248)
//   g2_b = β₂ + Σ wᵢ·B²ᵢ + s·δ₂
//   g_c  = s·g_a + r·g1_b − rs·δ₁ + Σ wᵢ·Lᵢ + Σ hᵢ·Hᵢ
// followed by into_affine() and ark-serialize's compressed encoding (rln/src/protocol/proof.rs:413-428).
//
// B200-first formulation: the bases never change (they come from the zkey), and a batch holds
// thousands of independent proofs.  Instead of running Pippenger per proof (bucket reduction would
// cost 25-40 % at n ≈ 6 K), every base gets a precomputed table T[base][window][d] = d·2^{c·window}·P
// for d = 1 … 2^{c−1} (signed digits), resident in HBM (c = 12: ≈ 90 GB of the 180 GB).  An MSM
// then is a pure stream of mixed additions of table entries — no doublings, no buckets — and a warp
// of 32 proofs walks the same (base, window) sub-table together.  Partial sums per (task, proof) are
// combined by a second kernel.  Additions are complete (∞, P = ±Q handled), because duplicate bases
// in a zkey cannot be excluded.
#include <cstdlib>
#include <vector>

#define ZK_FQ2_OUTLINE 1  // this TU only: Fq2 products call one shared out-of-line Fq multiplier (see fp.cuh mul_ni)
#include "device_api.hpp"
#include "fixed_base.cuh"
#include "glv.cuh"

namespace zk {

// ------------------------------------------------------------------------------------------- table construction
// wb[base*K + k] = 2^{c·k}·P_base (affine).  One thread per base.
template <class F>
__global__ void __launch_bounds__(64) k_window_bases(const Affine<F>* __restrict__ bases, u32 n, int c, int K, Affine<F>* __restrict__ wb) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> acc = XYZZ<F>::from_affine(bases[i]);
    wb[(size_t)i * K] = bases[i];
    for (int k = 1; k < K; k++) {
        for (int t = 0; t < c; t++) acc = acc.dbl();
        wb[(size_t)i * K + k] = acc.to_affine();
    }
}

// table[(base*K + k)*half + d−1] = d·wb[base*K + k], d = 1 … half.  One thread per (base, window);
// NB consecutive multiples are normalised with one shared inversion (Montgomery's trick).
template <class F, int NB>
__global__ void __launch_bounds__(64) k_fill_table(const Affine<F>* __restrict__ wb, size_t n_slots, u32 half, Affine<F>* __restrict__ table) {
    const size_t slot = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const Affine<F> P = wb[slot];
    Affine<F>* out = table + slot * half;
    XYZZ<F> acc = XYZZ<F>::from_affine(P);
    for (u32 d0 = 0; d0 < half; d0 += NB) {
        XYZZ<F> buf[NB];
        F pref[NB];
        F run = F::one();
#pragma unroll 1
        for (int t = 0; t < NB; t++) {
            buf[t] = acc;
            run = run * acc.ZZZ;
            pref[t] = run;
            acc.add_affine(P);
        }
        F inv = run.inv();
#pragma unroll 1
        for (int t = NB - 1; t >= 0; t--) {
            F zi = t ? inv * pref[t - 1] : inv;  // 1/ZZZ_t
            inv = inv * buf[t].ZZZ;
            F zz_inv = (zi * buf[t].ZZ).sqr();
            if (d0 + t < half) out[d0 + t] = {buf[t].X * zz_inv, buf[t].Y * zi};
        }
    }
}

void launch_build_table_g1(const G1Affine* d_bases, u32 n, int c, int K, G1Affine* d_table, cudaStream_t s) {
    if (!n) return;
    G1Affine* wb = nullptr;
    ZK_CUDA_CHECK(cudaMallocAsync((void**)&wb, sizeof(G1Affine) * (size_t)n * K, s));
    k_window_bases<Fq><<<(n + 63) / 64, 64, 0, s>>>(d_bases, n, c, K, wb);
    size_t slots = (size_t)n * K;
    k_fill_table<Fq, 16><<<(unsigned)((slots + 63) / 64), 64, 0, s>>>(wb, slots, 1u << (c - 1), d_table);
    ZK_CUDA_CHECK(cudaFreeAsync(wb, s));
}
void launch_build_table_g2(const G2Affine* d_bases, u32 n, int c, int K, G2Affine* d_table, cudaStream_t s) {
    if (!n) return;
    G2Affine* wb = nullptr;
    ZK_CUDA_CHECK(cudaMallocAsync((void**)&wb, sizeof(G2Affine) * (size_t)n * K, s));
    k_window_bases<Fq2><<<(n + 63) / 64, 64, 0, s>>>(d_bases, n, c, K, wb);
    size_t slots = (size_t)n * K;
    k_fill_table<Fq2, 8><<<(unsigned)((slots + 63) / 64), 64, 0, s>>>(wb, slots, 1u << (c - 1), d_table);
    ZK_CUDA_CHECK(cudaFreeAsync(wb, s));
}

// ------------------------------------------------------------------------------------------- GLV split (G1)
// BN254 G1 has the endomorphism φ(x, y) = (β·x, y) = [λ](x, y) with β³ = 1 in Fq, λ² + λ + 1 = 0 in Fr.  A scalar
// k < r splits as k ≡ k₁ + k₂·λ (mod r) with |k₁|, |k₂| < 2^128 (Babai rounding against the reduced lattice basis
// v₁ = (s, −N₁), v₂ = (N₂, s), s² + N₁·N₂ = r).  For the window tables this halves the number of windows per base:
// Σ k·P = Σ k₁·P + φ(Σ k₂·P), both sums over the SAME table, φ applied once per partial sum in k_msm_reduce.
// With c = 13 that is 2 × 10 additions per term instead of 22 at c = 12, in 10 % less table memory.
// (Same group element as ark-ec's msm_bigint, rln/src/partial_proof.rs:98-104 — the result is unique.)

__global__ void k_glv_split(const uint8_t* __restrict__ scalars, size_t n, uint8_t* __restrict__ out) {  // self-test: n × (16 B |k₁|, 16 B |k₂|, 2 sign bytes)
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 k[8], h[8];
    for (int w = 0; w < 8; w++) k[w] = reinterpret_cast<const u32*>(scalars + 32 * i)[w];
    for (int half = 0; half < 2; half++) {
        const bool neg = glv::split(k, half, h);
        for (int w = 0; w < 4; w++) reinterpret_cast<u32*>(out + 36 * i + 16 * half)[w] = h[w];
        out[36 * i + 32 + half] = neg ? 1 : 0;
    }
    out[36 * i + 34] = 0; out[36 * i + 35] = 0;
}
// self-test of glv_double_mul: per item P (64 B affine canonical), kp (32 B), Q (64 B), kq (32 B) → kp·P + kq·Q (64 B); P enters
// with non-trivial ZZ / ZZZ (2P − P) like the sums the assembly feeds it
__global__ void k_glv_double_mul(const uint8_t* __restrict__ in, size_t n, int use_q, uint8_t* __restrict__ out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* p = in + 192 * i;
    auto ld = [](const uint8_t* b) { u32 c[8]; for (int w = 0; w < 8; w++) c[w] = reinterpret_cast<const u32*>(b)[w]; return Fq::from_canonical(c); };
    const G1Affine Pa = {ld(p), ld(p + 32)}, Qa = {ld(p + 96), ld(p + 128)};
    u32 kp[8], kq[8];
    for (int w = 0; w < 8; w++) { kp[w] = reinterpret_cast<const u32*>(p + 64)[w]; kq[w] = reinterpret_cast<const u32*>(p + 160)[w]; }
    G1XYZZ P = G1XYZZ::from_affine(Pa).dbl();
    P.add(G1XYZZ::from_affine(Pa).neg());
    const G1Affine r = glv_double_mul(P, kp, G1XYZZ::from_affine(Qa), kq, use_q != 0).to_affine();
    u32 x[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (!r.is_inf()) { r.x.to_canonical(x); r.y.to_canonical(y); }
    for (int w = 0; w < 8; w++) { reinterpret_cast<u32*>(out + 64 * i)[w] = x[w]; reinterpret_cast<u32*>(out + 64 * i + 32)[w] = y[w]; }
}
void launch_glv_double_mul(const uint8_t* d_in, size_t n, int use_q, uint8_t* d_out, cudaStream_t s) {
    if (n) k_glv_double_mul<<<(unsigned)((n + 31) / 32), 32, 0, s>>>(d_in, n, use_q, d_out);
}
void launch_glv_split(const uint8_t* d_scalars, size_t n, uint8_t* d_out, cudaStream_t s) {
    if (n) k_glv_split<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_scalars, n, d_out);
}

// ------------------------------------------------------------------------------------------- accumulate

template <class F>
struct AccumArgs {
    const Fr* src[2];            // scalar matrices [row][B]: 0 = witness values, 1 = h
    const u32* row[4];           // per group: scalar row of each base
    const Affine<F>* table[4];   // per group
    u32 which[4];
    const MsmTask* tasks;
    XYZZ<F>* part;               // [task][B]
    u32 B;
    int c, K;
    int glv;                     // G1 only: tasks come in (k₁, k₂) pairs, MsmTask::half selects which
    // Small batches (B < 32, e.g. a single proof): a warp whose lanes are 32 consecutive proofs would run with 1 … 31 lanes idle
    // while still paying every IMAD.WIDE's 4 pipe cycles.  Packed, the lanes of a CTA are consecutive (task, proof) pairs, proof
    // fastest over 2^pack_log ≥ B slots: at B = 1 a warp carries 32 tasks (measured at B = 1, G1: 2.2 → see profiles/README.md).
    u32 packed, pack_log, n_tasks;
};

template <class F, bool PREFETCH, int MIN_BLOCKS, bool GLV = false>
__global__ void __launch_bounds__(128, MIN_BLOCKS) k_msm_accum(AccumArgs<F> a) {
    u32 j, task;
    if (a.packed) {
        const u32 slot = blockIdx.y * blockDim.x + threadIdx.x;
        j = slot & ((1u << a.pack_log) - 1);
        task = slot >> a.pack_log;
        if (task >= a.n_tasks) return;
    } else {
        j = blockIdx.x * blockDim.x + threadIdx.x;
        task = blockIdx.y;
    }
    if (j >= a.B) return;
    const MsmTask t = a.tasks[task];
    const u32* __restrict__ rows = a.row[t.group];
    const Affine<F>* __restrict__ table = a.table[t.group];
    const Fr* __restrict__ src = a.src[a.which[t.group]];
    const u32 half = 1u << (a.c - 1);
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (u32 b = t.lo; b < t.hi; b++) {
        u32 s[8];
        ld_fp(src + (size_t)rows[b] * a.B + j).to_canonical(s);
        bool flip = false;       // GLV: the half-scalar is negative, every digit changes sign
        if (GLV) {
            u32 h[8];
            flip = glv::split(s, t.half, h);
#pragma unroll
            for (int i = 0; i < 8; i++) s[i] = h[i];
        }
        u32 any = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) any |= s[i];
        if (!any) continue;
        const Affine<F>* tb = table + (size_t)b * a.K * half;
        u32 carry = 0;
        if (!PREFETCH) {
            for (int k = 0; k < a.K; k++) {
                const int d = window_digit(s, k, a.c, carry);
                if (d == 0) continue;
                Affine<F> pt = ld_point<F>(tb + (size_t)k * half + ((d < 0 ? -d : d) - 1));
                if ((d < 0) != flip) pt.y = pt.y.neg();
                acc.add_affine(pt);
            }
            continue;
        }
        // software pipeline: fetch the table entry of window k+1 while adding that of window k
        int d = window_digit(s, 0, a.c, carry);
        Affine<F> nxt;
        bool nxt_valid = d != 0;
        if (nxt_valid) {
            nxt = ld_point<F>(tb + ((d < 0 ? -d : d) - 1));
            if ((d < 0) != flip) nxt.y = nxt.y.neg();
        }
        for (int k = 0; k < a.K; k++) {
            Affine<F> cur = nxt;
            const bool cur_valid = nxt_valid;
            nxt_valid = false;
            if (k + 1 < a.K) {
                d = window_digit(s, k + 1, a.c, carry);
                nxt_valid = d != 0;
                if (nxt_valid) {
                    nxt = ld_point<F>(tb + (size_t)(k + 1) * half + ((d < 0 ? -d : d) - 1));
                    if ((d < 0) != flip) nxt.y = nxt.y.neg();
                }
            }
            if (cur_valid) acc.add_affine(cur);
        }
    }
    a.part[(size_t)task * a.B + j] = acc;
}


// partial sum of a task; the k₂ half of a GLV pair goes through φ: (X, Y, ZZ, ZZZ) ↦ (β·X, Y, ZZ, ZZZ)
__device__ __forceinline__ G1XYZZ load_partial(const G1XYZZ* p, u32 half) {
    G1XYZZ v = *p;
    if (half) v.X = v.X * glv::beta();
    return v;
}
// on G2 (the sextic twist over Fq2) the same λ acts as (x, y) ↦ (β²·x, y)
__device__ __forceinline__ G2XYZZ load_partial(const G2XYZZ* p, u32 half) {
    G2XYZZ v = *p;
    if (half) v.X = v.X.scale(glv::beta2());
    return v;
}

__device__ __forceinline__ Fq shfl_xor_coord(const Fq& v, int m) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, v.l[i], m);
    return r;
}
__device__ __forceinline__ Fq2 shfl_xor_coord(const Fq2& v, int m) { return {shfl_xor_coord(v.a, m), shfl_xor_coord(v.b, m)}; }
// sum[group][j] = Σ_{tasks of group} part[task][j].  Large batches: a (proof, group) pair has ≈ 110 partial sums, a chain of
// dependent additions that one thread per proof walked alone (G1 1.27 ms, G2 2.14 ms at 4 096 proofs with 16 K threads on the
// chip); REDUCE_LANES adjacent lanes now take every REDUCE_LANES-th partial of the pair and meet in a shuffle tree.
static const u32 REDUCE_LANES = 8;
template <class F>
__global__ void __launch_bounds__(128) k_msm_reduce(const XYZZ<F>* __restrict__ part, const MsmTask* __restrict__ tasks, u32 n_tasks,
                                                    u32 B, XYZZ<F>* __restrict__ sum) {
    const u32 gid = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 j = gid / REDUCE_LANES, q = gid % REDUCE_LANES;
    const u32 g = blockIdx.y;
    const bool live = j < B;                       // B is a multiple of 32 / REDUCE_LANES on this path or the last warp is ragged: no early return before the shuffles
    XYZZ<F> acc = XYZZ<F>::infinity();
    // the tasks of a group are contiguous (msm_make_tasks): [lo, hi); the lanes of a pair step through it side by side, so that the
    // additions of a warp fall into the same iterations
    u32 lo = 0, hi = 0;
    for (u32 t = 0; t < n_tasks; t++) {
        const u32 tg = tasks[t].group;
        if (tg < g) lo = t + 1;
        if (tg <= g) hi = t + 1;
    }
    if (live)
        for (u32 t = lo + q; t < hi; t += REDUCE_LANES) acc.add(load_partial(part + (size_t)t * B + j, tasks[t].half));
#pragma unroll
    for (int m = 1; m < (int)REDUCE_LANES; m <<= 1) {
        XYZZ<F> o = {shfl_xor_coord(acc.X, m), shfl_xor_coord(acc.Y, m), shfl_xor_coord(acc.ZZ, m), shfl_xor_coord(acc.ZZZ, m)};
        acc.add(o);
    }
    if (live && q == 0) sum[(size_t)g * B + j] = acc;
}

// Small batches (a single proof through ffi_generate_rln_proof): thousands of partials per (group, proof) would be
// summed by one thread above.  Here one CTA owns a (proof, group) pair: 128 threads stride over the tasks, then a
// shared-memory tree combines them.
template <class F>
__global__ void __launch_bounds__(128) k_msm_reduce_small(const XYZZ<F>* __restrict__ part, const MsmTask* __restrict__ tasks, u32 n_tasks,
                                                          u32 B, XYZZ<F>* __restrict__ sum) {
    __shared__ XYZZ<F> sh[128];
    const u32 j = blockIdx.x, g = blockIdx.y;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (u32 t = threadIdx.x; t < n_tasks; t += 128)
        if (tasks[t].group == g) acc.add(load_partial(part + (size_t)t * B + j, tasks[t].half));
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (u32 w = 64; w >= 1; w >>= 1) {
        if (threadIdx.x < w) {
            XYZZ<F> a = sh[threadIdx.x];
            a.add(sh[threadIdx.x + w]);
            sh[threadIdx.x] = a;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sum[(size_t)g * B + j] = sh[0];
}

// ------------------------------------------------------------------------------------------- assembly
__device__ __forceinline__ void load_scalar_bytes(const uint8_t* p, u32* out) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
    u32 m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = FrCfg::p(i);
    while (Fr::raw_cmp(out, m) >= 0) Fr::raw_sub(out, out, m);
}
__device__ __forceinline__ bool fq_is_larger_half(const u32* y) {  // y > q − y  (y canonical, non-zero)
    u32 q[8], n[8];
#pragma unroll
    for (int i = 0; i < 8; i++) q[i] = FqCfg::p(i);
    Fq::raw_sub(n, q, y);
    return Fq::raw_cmp(y, n) > 0;
}
__device__ __forceinline__ void store_words(uint8_t* p, const u32* w) {
    u32* o = reinterpret_cast<u32*>(p);
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = w[i];
}
// ark-serialize 0.5 compressed short-Weierstrass point: x little-endian, bit 7 of the last byte =
// "y is the larger of {y, −y}", bit 6 = infinity.
__device__ void compress_g1(const G1Affine& p, uint8_t* out, uint8_t* affine_out) {
    u32 x[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool inf = p.is_inf();
    if (!inf) {
        p.x.to_canonical(x);
        p.y.to_canonical(y);
    }
    if (affine_out) {
        store_words(affine_out, x);
        store_words(affine_out + 32, y);
        if (inf) affine_out[63] = 0x40;
    }
    if (inf) x[7] |= 0x40000000u;
    else if (fq_is_larger_half(y)) x[7] |= 0x80000000u;
    store_words(out, x);
}
__device__ void compress_g2(const G2Affine& p, uint8_t* out, uint8_t* affine_out) {
    u32 x0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, x1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, y1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool inf = p.is_inf();
    if (!inf) {
        p.x.a.to_canonical(x0);
        p.x.b.to_canonical(x1);
        p.y.a.to_canonical(y0);
        p.y.b.to_canonical(y1);
    }
    if (affine_out) {
        store_words(affine_out, x0);
        store_words(affine_out + 32, x1);
        store_words(affine_out + 64, y0);
        store_words(affine_out + 96, y1);
        if (inf) affine_out[127] = 0x40;
    }
    if (inf) x1[7] |= 0x40000000u;
    else {
        // Fq2 ordering: compare c1 first, then c0 (y vs −y)
        bool larger;
        bool c1zero = true;
        for (int i = 0; i < 8; i++) c1zero = c1zero && y1[i] == 0;
        if (!c1zero) larger = fq_is_larger_half(y1);
        else larger = fq_is_larger_half(y0);
        if (larger) x1[7] |= 0x80000000u;
    }
    store_words(out, x0);
    store_words(out + 32, x1);
}

__device__ __forceinline__ Fq load_fq_canonical(const uint8_t* p, bool mask_flags) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    u32 c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    if (mask_flags) c[7] &= 0x3fffffffu;
    return Fq::from_canonical(c);
}
__device__ __forceinline__ G1Affine load_affine_g1(const uint8_t* p) {  // x|y canonical, 0x40 in the last byte = infinity
    if (p[63] & 0x40) return G1Affine::infinity();
    return {load_fq_canonical(p, false), load_fq_canonical(p + 32, true)};
}
__device__ __forceinline__ G2Affine load_affine_g2(const uint8_t* p) {
    if (p[127] & 0x40) return G2Affine::infinity();
    return {{load_fq_canonical(p, false), load_fq_canonical(p + 32, false)}, {load_fq_canonical(p + 64, false), load_fq_canonical(p + 96, true)}};
}

// partial proof points (partial_proof.rs:159-170): π_a = α₁ + ΣA, ρ = β₁ + ΣB₁, π_b = β₂ + ΣB₂, π_c = ΣL over the known wires
__global__ void __launch_bounds__(64) k_partial_g1(ProverKeyDev pk, const G1XYZZ* __restrict__ sum, u32 B, uint8_t* __restrict__ out_affine,
                                                   uint8_t* __restrict__ out_comp) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    G1XYZZ a = sum[0 * (size_t)B + j], b = sum[1 * (size_t)B + j], c = sum[2 * (size_t)B + j];
    a.add_affine(pk.alpha_g1);
    b.add_affine(pk.beta_g1);
    uint8_t* af = out_affine + 320 * (size_t)j;
    uint8_t* cp = out_comp + 160 * (size_t)j;
    compress_g1(a.to_affine(), cp, af);
    compress_g1(b.to_affine(), cp + 32, af + 64);
    compress_g1(c.to_affine(), cp + 128, af + 256);
}
__global__ void __launch_bounds__(64) k_partial_g2(ProverKeyDev pk, const G2XYZZ* __restrict__ sum, u32 B, uint8_t* __restrict__ out_affine,
                                                   uint8_t* __restrict__ out_comp) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    G2XYZZ b = sum[j];
    b.add_affine(pk.beta_g2);
    compress_g2(b.to_affine(), out_comp + 160 * (size_t)j + 64, out_affine + 320 * (size_t)j + 128);
}

__global__ void __launch_bounds__(64) k_assemble_g1(ProverKeyDev pk, const G1Affine* __restrict__ dtab, int c, int K,
                                                    const G1XYZZ* __restrict__ sum, u32 B, const uint8_t* __restrict__ rs,
                                                    const uint8_t* __restrict__ partial, uint8_t* __restrict__ proofs,
                                                    uint8_t* __restrict__ affine) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    G1Affine base_a = pk.alpha_g1, base_b = pk.beta_g1, base_c = G1Affine::infinity();
    if (partial) {  // finish phase: the partial proof already contains α₁ / β₁ and the known part of L
        const uint8_t* pp = partial + 320 * (size_t)j;
        base_a = load_affine_g1(pp);
        base_b = load_affine_g1(pp + 64);
        base_c = load_affine_g1(pp + 256);
    }
    u32 r[8], s[8], rsv[8];
    load_scalar_bytes(rs + 64 * (size_t)j, r);
    load_scalar_bytes(rs + 64 * (size_t)j + 32, s);
    (Fr::from_canonical(r) * Fr::from_canonical(s)).to_canonical(rsv);
    u32 rnz = 0;
    for (int i = 0; i < 8; i++) rnz |= r[i];
    // g_a
    G1XYZZ g_a = sum[0 * (size_t)B + j];
    if (!base_a.is_inf()) g_a.add_affine(base_a);
    g_a.add(fixed_base_mul<Fq>(dtab, c, K, r));
    // g_c = s·g_a + r·g1_b − rs·δ₁ + L + H
    G1XYZZ g1_b = G1XYZZ::infinity();
    if (rnz) {
        g1_b = sum[1 * (size_t)B + j];
        if (!base_b.is_inf()) g1_b.add_affine(base_b);
        g1_b.add(fixed_base_mul<Fq>(dtab, c, K, s));
    }
    G1XYZZ g_c = glv_double_mul(g_a, s, g1_b, r, rnz != 0);
    g_c.add(fixed_base_mul<Fq>(dtab, c, K, rsv).neg());
    g_c.add(sum[2 * (size_t)B + j]);
    g_c.add(sum[3 * (size_t)B + j]);
    if (!base_c.is_inf()) g_c.add_affine(base_c);
    uint8_t* o = proofs + 128 * (size_t)j;
    uint8_t* af = affine ? affine + 256 * (size_t)j : nullptr;
    compress_g1(g_a.to_affine(), o, af);
    compress_g1(g_c.to_affine(), o + 96, af ? af + 192 : nullptr);
}
__global__ void __launch_bounds__(64) k_assemble_g2(ProverKeyDev pk, const G2Affine* __restrict__ dtab, int c, int K,
                                                    const G2XYZZ* __restrict__ sum, u32 B, const uint8_t* __restrict__ rs,
                                                    const uint8_t* __restrict__ partial, uint8_t* __restrict__ proofs,
                                                    uint8_t* __restrict__ affine) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B) return;
    u32 s[8];
    load_scalar_bytes(rs + 64 * (size_t)j + 32, s);
    G2XYZZ g2_b = sum[j];
    G2Affine base_b = partial ? load_affine_g2(partial + 320 * (size_t)j + 128) : pk.beta_g2;
    if (!base_b.is_inf()) g2_b.add_affine(base_b);
    g2_b.add(fixed_base_mul<Fq2>(dtab, c, K, s));
    compress_g2(g2_b.to_affine(), proofs + 128 * (size_t)j + 32, affine ? affine + 256 * (size_t)j + 64 : nullptr);
}

// The same assembly for a handful of proofs (a single ffi_generate_rln_proof call): one CTA per proof, the independent chains on
// four warps — four schedulers — instead of one thread doing them one after the other (2.0 ms of a single proof's 8 ms).  With
// P = ΣA + α₁ and Q = ΣB₁ + β₁,   s·g_a + r·g1_b − rs·δ₁ = s·P + r·Q + rs·δ₁,   so the Straus double multiplication does not wait for
// the r·δ₁ / s·δ₁ terms:  warp 0: π_c = s·P + r·Q + (warp 2's sum), warp 1: π_a = P + r·δ₁, warp 2: rs·δ₁ + ΣL + ΣH, warp 3: π_b.
// Same group elements as k_assemble_g1 / _g2, hence the same bytes.
__global__ void __launch_bounds__(160) k_assemble_small(ProverKeyDev pk, const G1Affine* __restrict__ dtab, int c, int K,
                                                        const G2Affine* __restrict__ dtab2, int c2, int K2, const G1XYZZ* __restrict__ sum,
                                                        const G2XYZZ* __restrict__ sum2, u32 B, const uint8_t* __restrict__ rs,
                                                        const uint8_t* __restrict__ partial, uint8_t* __restrict__ proofs,
                                                        uint8_t* __restrict__ affine, const G1XYZZ* __restrict__ fold,
                                                        const G1Affine* __restrict__ atab, const G1Affine* __restrict__ btab) {
    __shared__ G1XYZZ s_tail, s_beta;
    if (fold) {
        // Folded form (full proofs only): s·P + r·Q = s·α₁ + r·β₁ + Σ(s·zᵢ)Aᵢ + Σ(r·zᵢ)B₁ᵢ, the two sums being fold[0], fold[1] — table
        // sums like every other, so no variable-base multiplication is left (the Straus run was ≈ 3 000 dependent products).
        // warp 0: π_c, warp 1: π_a, warp 2: rs·δ₁ + ΣL + ΣH, warp 3: π_b, warp 4: r·β₁.
        const u32 j = blockIdx.x, warp = threadIdx.x >> 5;
        const bool lead = (threadIdx.x & 31) == 0;
        u32 r[8], s[8];
        load_scalar_bytes(rs + 64 * (size_t)j, r);
        load_scalar_bytes(rs + 64 * (size_t)j + 32, s);
        uint8_t* o = proofs + 128 * (size_t)j;
        uint8_t* af = affine ? affine + 256 * (size_t)j : nullptr;
        if (warp == 3) {
            if (!lead) return;
            G2XYZZ g2_b = sum2[j];
            g2_b.add_affine(pk.beta_g2);
            g2_b.add(fixed_base_mul<Fq2>(dtab2, c2, K2, s));
            compress_g2(g2_b.to_affine(), o + 32, af ? af + 64 : nullptr);
            return;
        }
        if (warp == 1) {
            if (!lead) return;
            G1XYZZ P = sum[0 * (size_t)B + j];
            P.add_affine(pk.alpha_g1);
            P.add(fixed_base_mul<Fq>(dtab, c, K, r));
            compress_g1(P.to_affine(), o, af);
            return;
        }
        if (warp == 2 || warp == 4) {
            if (lead) {
                if (warp == 2) {
                    u32 rsv[8];
                    (Fr::from_canonical(r) * Fr::from_canonical(s)).to_canonical(rsv);
                    G1XYZZ t = fixed_base_mul<Fq>(dtab, c, K, rsv);
                    t.add(sum[2 * (size_t)B + j]);
                    t.add(sum[3 * (size_t)B + j]);
                    s_tail = t;
                } else {
                    G1XYZZ t = fixed_base_mul<Fq>(btab, c, K, r);
                    t.add(fold[1 * (size_t)B + j]);
                    s_beta = t;
                }
            }
            __syncwarp();
            asm volatile("bar.sync 1, 96;" ::: "memory");   // with warp 0
            return;
        }
        G1XYZZ g_c = G1XYZZ::infinity();
        if (lead && warp == 0) {
            g_c = fixed_base_mul<Fq>(atab, c, K, s);
            g_c.add(fold[0 * (size_t)B + j]);
        }
        __syncwarp();
        asm volatile("bar.sync 1, 96;" ::: "memory");
        if (lead) {
            g_c.add(s_beta);
            g_c.add(s_tail);
            compress_g1(g_c.to_affine(), o + 96, af ? af + 192 : nullptr);
        }
        return;
    }
    if (threadIdx.x >= 128) return;
    const u32 j = blockIdx.x, warp = threadIdx.x >> 5;
    const bool lead = (threadIdx.x & 31) == 0;   // one lane per warp works; the others stay for the warp-wide barrier
    u32 r[8], s[8];
    load_scalar_bytes(rs + 64 * (size_t)j, r);
    load_scalar_bytes(rs + 64 * (size_t)j + 32, s);
    uint8_t* o = proofs + 128 * (size_t)j;
    uint8_t* af = affine ? affine + 256 * (size_t)j : nullptr;
    const uint8_t* pp = partial ? partial + 320 * (size_t)j : nullptr;
    if (warp == 3) {
        if (!lead) return;
        G2XYZZ g2_b = sum2[j];
        const G2Affine base_b = pp ? load_affine_g2(pp + 128) : pk.beta_g2;
        if (!base_b.is_inf()) g2_b.add_affine(base_b);
        g2_b.add(fixed_base_mul<Fq2>(dtab2, c2, K2, s));
        compress_g2(g2_b.to_affine(), o + 32, af ? af + 64 : nullptr);
        return;
    }
    if (warp == 2) {
        if (lead) {
            u32 rsv[8];
            (Fr::from_canonical(r) * Fr::from_canonical(s)).to_canonical(rsv);
            G1XYZZ t = fixed_base_mul<Fq>(dtab, c, K, rsv);
            t.add(sum[2 * (size_t)B + j]);
            t.add(sum[3 * (size_t)B + j]);
            if (pp) { const G1Affine base_c = load_affine_g1(pp + 256); if (!base_c.is_inf()) t.add_affine(base_c); }
            s_tail = t;
        }
        __syncwarp();
        asm volatile("bar.sync 1, 64;" ::: "memory");   // with warp 0
        return;
    }
    if (warp == 1) {
        if (!lead) return;
        G1XYZZ P = sum[0 * (size_t)B + j];
        const G1Affine base_a = pp ? load_affine_g1(pp) : pk.alpha_g1;
        if (!base_a.is_inf()) P.add_affine(base_a);
        P.add(fixed_base_mul<Fq>(dtab, c, K, r));
        compress_g1(P.to_affine(), o, af);
        return;
    }
    G1XYZZ g_c = G1XYZZ::infinity();
    if (lead) {
        G1XYZZ P = sum[0 * (size_t)B + j];
        const G1Affine base_a = pp ? load_affine_g1(pp) : pk.alpha_g1;
        if (!base_a.is_inf()) P.add_affine(base_a);
        u32 rnz = 0;
        for (int i = 0; i < 8; i++) rnz |= r[i];
        G1XYZZ Q = G1XYZZ::infinity();
        if (rnz) {
            Q = sum[1 * (size_t)B + j];
            const G1Affine base_b = pp ? load_affine_g1(pp + 64) : pk.beta_g1;
            if (!base_b.is_inf()) Q.add_affine(base_b);
        }
        g_c = glv_double_mul(P, s, Q, r, rnz != 0);
    }
    __syncwarp();
    asm volatile("bar.sync 1, 64;" ::: "memory");   // warp 2's sum is in shared memory
    if (lead) {
        g_c.add(s_tail);
        compress_g1(g_c.to_affine(), o + 96, af ? af + 192 : nullptr);
    }
}

__global__ void __launch_bounds__(128) k_scale_vals(const Fr* __restrict__ vals, const uint8_t* __restrict__ rs, u32 n_rows, u32 B,
                                                    Fr* __restrict__ out_s, Fr* __restrict__ out_r) {
    const size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (t >= (size_t)n_rows * B) return;
    const u32 j = (u32)(t % B);
    u32 r[8], s[8];
    load_scalar_bytes(rs + 64 * (size_t)j, r);
    load_scalar_bytes(rs + 64 * (size_t)j + 32, s);
    const Fr v = ld_fp(vals + t);
    st_fp(out_s + t, v * Fr::from_canonical(s));
    st_fp(out_r + t, v * Fr::from_canonical(r));
}
void launch_scale_vals(const Fr* d_vals, const uint8_t* d_rs, u32 n_rows, u32 B, Fr* d_out_s, Fr* d_out_r, cudaStream_t s) {
    const size_t n = (size_t)n_rows * B;
    k_scale_vals<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(d_vals, d_rs, n_rows, B, d_out_s, d_out_r);
}

// ------------------------------------------------------------------------------------------- host orchestration
// bases per task.  A thread walks `chunk` bases for one proof, a CTA carries 128 proofs, and the CTAs of one launch run in
// waves of (SMs × resident CTAs): long CTAs in few waves leave part of the chip idle at the end of the launch, short CTAs in
// many waves cost more partial sums for k_msm_reduce.  Measured at batch 4 096: G1 10 → 24 waves −1.5 % (reduce +0.75 ms);
// G2 is flat between 8 and 24 waves because its reduce (Fq2 additions) grows as fast as the tail shrinks.
// Bases per task: the smallest chunk whose task count fits the target number of "waves" of 148 × (4 | 2) resident CTAs (24 for
// G1, 12 for G2), so that the launch does not end with a nearly empty extra wave.  Measured at batch 4 096 (round 2): G1 452 tasks =
// 24.4 waves → 442 = 23.9: 285.0 → 283.8 ms; G2 112 = 12.1 → 110 = 11.9: 131.7 → 131.7 ms — CTAs of a launch drift apart, so there
// is hardly a wave structure left to fit by the end; kept because it is never worse.
static u32 pick_chunk(const u32* group_sizes, int n_groups, u32 halves, u32 B, bool g2) {
    const u64 ctas_per_task = (B + 127) / 128;
    const u64 per_wave = 148ull * (g2 ? 2 : 4);
    u64 max_tasks = (g2 ? 12 : 24) * per_wave / ctas_per_task;
    if (max_tasks < halves) max_tasks = halves;
    for (u32 chunk = 4; chunk < 256; chunk++) {
        u64 tasks = 0;
        for (int i = 0; i < n_groups; i++) tasks += (u64)((group_sizes[i] + chunk - 1) / chunk) * halves;
        if (tasks <= max_tasks) return chunk;
    }
    return 256;
}
std::vector<MsmTask> msm_make_tasks(const FixedMsmPlan& plan, u32 B, bool g2, int phase) {
    const MsmGroupDev* g = g2 ? &plan.g2 : plan.g1;
    const int n_groups = g2 ? 1 : 4;
    auto range = [&](int i, u32& lo, u32& hi) {
        lo = 0;
        hi = g[i].n_bases;
        if (phase == MSM_KNOWN) hi = g[i].n_known;      // H has n_known = 0: it belongs to the finish phase only
        if (phase == MSM_UNKNOWN) lo = g[i].n_known;
    };
    u32 sizes[4] = {0, 0, 0, 0};
    for (int i = 0; i < n_groups; i++) { u32 lo, hi; range(i, lo, hi); sizes[i] = hi - lo; }
    const bool glv = plan.glv != 0;   // both groups: G1 through β, G2 through β²
    const u32 chunk = pick_chunk(sizes, n_groups, glv ? 2 : 1, B, g2);
    std::vector<MsmTask> tasks;
    for (int i = 0; i < n_groups; i++) {
        u32 lo, hi;
        range(i, lo, hi);
        for (u32 b = lo; b < hi; b += chunk)
            for (u32 h = 0; h < (glv ? 2u : 1u); h++) tasks.push_back({(u32)i, b, b + chunk < hi ? b + chunk : hi, h});
    }
    return tasks;
}


// grid of an accumulate launch over `count` tasks: one CTA row per task for full warps of proofs, packed (task, proof) lanes below 32
template <class F>
static dim3 accum_grid(AccumArgs<F>& a, u32 B, u32 bx, u32 count) {
    a.n_tasks = count;
    a.packed = B < 32 ? 1 : 0;
    a.pack_log = 0;
    if (!a.packed) return dim3((B + bx - 1) / bx, count);
    while ((1u << a.pack_log) < B) a.pack_log++;
    const u64 slots = (u64)count << a.pack_log;
    return dim3(1, (unsigned)((slots + bx - 1) / bx));
}
// thread-per-proof loop over a group's partial sums, or one CTA per (proof, group) with a shared-memory tree?  The loop costs
// ≈ 2.8 µs per partial (G1), the tree ≈ 2.2 ms whatever B is (the number of partials, tasks × B, is roughly constant by
// construction of pick_chunk): measured B = 256: loop 21.1 ms (7 890 tasks) / B = 4 096: loop 1.27 ms, tree 2.15 ms (447 tasks).
static bool reduce_by_tree(u32 B, u32 n_tasks, bool g2) { return B < 64 || n_tasks > (g2 ? 400u : 700u); }

void launch_msm_sums(const FixedMsmPlan& plan, const Fr* d_vals, const Fr* d_h, u32 B, MsmWorkspace& ws, cudaStream_t s) {
    const u32 bx = B >= 128 || B < 32 ? 128 : 32;
    // a handful of proofs: none of these launches fills the chip, so the G2 sum runs beside the G1 sums on the caller's second stream
    const bool g2_aside = B <= 32 && ws.side && ws.side_fork && ws.side_join;
    cudaStream_t s2 = s;
    if (g2_aside) {
        ZK_CUDA_CHECK(cudaEventRecord(ws.side_fork, s));
        ZK_CUDA_CHECK(cudaStreamWaitEvent(ws.side, ws.side_fork, 0));
        s2 = ws.side;
    }
    {   // G1: A, B1, L, H
        AccumArgs<Fq> a;
        a.src[0] = d_vals; a.src[1] = d_h;
        for (int i = 0; i < 4; i++) { a.row[i] = plan.g1[i].row; a.table[i] = (const G1Affine*)plan.g1[i].table; a.which[i] = plan.g1[i].which_src; }
        a.tasks = ws.tasks_g1; a.part = ws.part_g1; a.B = B; a.c = plan.c; a.K = plan.K; a.glv = plan.glv;
        if (ws.ev) cudaEventRecord(ws.ev[0], s);
        auto launch_g1 = [&](u32 first, u32 count) {
            if (!count) return;
            AccumArgs<Fq> b = a;
            b.tasks = ws.tasks_g1 + first;
            b.part = ws.part_g1 + (size_t)first * B;
            const dim3 grid = accum_grid(b, B, bx, count);
            // prefetch of the next table entry on, 4 CTAs/SM: measured 1.8 % faster than 3 CTAs/SM, the other variants slower still
            // (round 1 A/B)
            if (plan.glv) k_msm_accum<Fq, true, 4, true><<<grid, bx, 0, s>>>(b);
            else k_msm_accum<Fq, true, 3><<<grid, bx, 0, s>>>(b);
        };
        launch_g1(0, ws.n_tasks_g1);
        if (ws.fold_s && ws.n_tasks_ab) {   // the same A and B₁ tasks once more over s·z and r·z (folded assembly, few proofs)
            AccumArgs<Fq> b = a;
            b.src[0] = ws.fold_s; b.src[1] = ws.fold_r;
            b.which[0] = 0; b.which[1] = 1;
            b.part = ws.fold_part;
            const dim3 grid = accum_grid(b, B, bx, ws.n_tasks_ab);
            if (plan.glv) k_msm_accum<Fq, true, 4, true><<<grid, bx, 0, s>>>(b);
            else k_msm_accum<Fq, true, 3><<<grid, bx, 0, s>>>(b);
        }
        if (ws.ev) cudaEventRecord(ws.ev[1], s);
        // one CTA per (proof, group) with a shared-memory tree whenever there are many partials per proof: a thread-per-proof loop
        // over hundreds of partials leaves the chip idle (4 096 threads)
        // of hundreds of partials leaves the chip idle; at batch 4 096 the tree is slower (G1 1.27 → 2.15 ms, round 1 A/B)
        const u32 rx = B >= 128 ? 128 : 32;
        if (reduce_by_tree(B, ws.n_tasks_g1, false)) k_msm_reduce_small<Fq><<<dim3(B, 4), 128, 0, s>>>(ws.part_g1, ws.tasks_g1, ws.n_tasks_g1, B, ws.sum_g1);
        else k_msm_reduce<Fq><<<dim3((B * REDUCE_LANES + rx - 1) / rx, 4), rx, 0, s>>>(ws.part_g1, ws.tasks_g1, ws.n_tasks_g1, B, ws.sum_g1);
        if (ws.fold_s && ws.n_tasks_ab) k_msm_reduce_small<Fq><<<dim3(B, 2), 128, 0, s>>>(ws.fold_part, ws.tasks_g1, ws.n_tasks_ab, B, ws.fold_sum);
        if (ws.ev) cudaEventRecord(ws.ev[2], s);
    }
    {   // G2: B2
        cudaStream_t s = s2;
        AccumArgs<Fq2> a;
        a.src[0] = d_vals; a.src[1] = d_h;
        for (int i = 0; i < 4; i++) { a.row[i] = plan.g2.row; a.table[i] = (const G2Affine*)plan.g2.table; a.which[i] = plan.g2.which_src; }
        a.tasks = ws.tasks_g2; a.part = ws.part_g2; a.B = B; a.c = plan.c2; a.K = plan.K2; a.glv = plan.glv;
        if (ws.n_tasks_g2) {
            const dim3 grid = accum_grid(a, B, bx, ws.n_tasks_g2);
            // no prefetch, 2 CTAs/SM: the G2 kernel is register-bound (prefetch −1.7 %, 3 CTAs/SM spills: −9 %; round 1 A/B)
            // Round 2 tried the other way to a third CTA per SM: X, Y, ZZ, ZZZ of every thread in shared memory, fetched where the
            // mixed addition uses them (168 registers, 88 B of spills, 3 CTAs/SM): 131.3 → 144.4 ms at batch 4 096 — the extra
            // shared-memory round trips sit on the dependent chain of every addition.  Measured and removed.
            if (plan.glv) k_msm_accum<Fq2, false, 2, true><<<grid, bx, 0, s>>>(a);
            else k_msm_accum<Fq2, false, 2><<<grid, bx, 0, s>>>(a);
        }
        if (ws.ev && !g2_aside) cudaEventRecord(ws.ev[3], s);
        const u32 rx = B >= 128 ? 128 : 32;
        if (reduce_by_tree(B, ws.n_tasks_g2, true)) k_msm_reduce_small<Fq2><<<dim3(B, 1), 128, 0, s>>>(ws.part_g2, ws.tasks_g2, ws.n_tasks_g2, B, ws.sum_g2);
        else k_msm_reduce<Fq2><<<dim3((B * REDUCE_LANES + rx - 1) / rx, 1), rx, 0, s>>>(ws.part_g2, ws.tasks_g2, ws.n_tasks_g2, B, ws.sum_g2);
        if (ws.ev && !g2_aside) cudaEventRecord(ws.ev[4], s);
    }
    if (g2_aside) {
        ZK_CUDA_CHECK(cudaEventRecord(ws.side_join, s2));
        ZK_CUDA_CHECK(cudaStreamWaitEvent(s, ws.side_join, 0));
        if (ws.ev) { cudaEventRecord(ws.ev[3], s); cudaEventRecord(ws.ev[4], s); }   // the G2 stages are hidden behind the G1 ones
    }
}

void launch_assemble(const FixedMsmPlan& plan, const ProverKeyDev& pk, u32 B, const uint8_t* d_rs, MsmWorkspace& ws,
                     const uint8_t* d_partial, uint8_t* d_proofs_out, uint8_t* d_proofs_affine, cudaStream_t s) {
    // π_b (G2) does not depend on π_a / π_c (G1) and both are latency-bound chains of one thread per proof: side by side when the
    // caller lends a second stream (they write disjoint bytes of every proof)
    if (B <= 32) {   // few proofs: latency is what counts — one CTA per proof, the chains on four warps, one launch
        const bool fold = ws.fold_sum && !d_partial;
        k_assemble_small<<<B, fold ? 160 : 128, 0, s>>>(pk, plan.delta1_table, plan.cd, plan.Kd, plan.delta2_table, plan.cd2, plan.Kd2, ws.sum_g1, ws.sum_g2,
                                                        B, d_rs, d_partial, d_proofs_out, d_proofs_affine, fold ? ws.fold_sum : nullptr,
                                                        plan.alpha1_table, plan.beta1_table);
    } else if (ws.side) {
        ZK_CUDA_CHECK(cudaEventRecord(ws.side_fork, s));
        ZK_CUDA_CHECK(cudaStreamWaitEvent(ws.side, ws.side_fork, 0));
        k_assemble_g2<<<(B + 63) / 64, 64, 0, ws.side>>>(pk, plan.delta2_table, plan.cd2, plan.Kd2, ws.sum_g2, B, d_rs, d_partial, d_proofs_out, d_proofs_affine);
        ZK_CUDA_CHECK(cudaEventRecord(ws.side_join, ws.side));
        k_assemble_g1<<<(B + 63) / 64, 64, 0, s>>>(pk, plan.delta1_table, plan.cd, plan.Kd, ws.sum_g1, B, d_rs, d_partial, d_proofs_out, d_proofs_affine);
        ZK_CUDA_CHECK(cudaStreamWaitEvent(s, ws.side_join, 0));
    } else {
        k_assemble_g1<<<(B + 63) / 64, 64, 0, s>>>(pk, plan.delta1_table, plan.cd, plan.Kd, ws.sum_g1, B, d_rs, d_partial, d_proofs_out, d_proofs_affine);
        k_assemble_g2<<<(B + 63) / 64, 64, 0, s>>>(pk, plan.delta2_table, plan.cd2, plan.Kd2, ws.sum_g2, B, d_rs, d_partial, d_proofs_out, d_proofs_affine);
    }
    if (ws.ev) cudaEventRecord(ws.ev[5], s);
}

void launch_partial_out(const ProverKeyDev& pk, u32 B, MsmWorkspace& ws, uint8_t* d_out_affine, uint8_t* d_out_compressed, cudaStream_t s) {
    k_partial_g1<<<(B + 63) / 64, 64, 0, s>>>(pk, ws.sum_g1, B, d_out_affine, d_out_compressed);
    k_partial_g2<<<(B + 63) / 64, 64, 0, s>>>(pk, ws.sum_g2, B, d_out_affine, d_out_compressed);
}

}  // namespace zk
