// Batched witness generation and QAP witness-map kernels (sm_100a).
//
// Path covered:
//   rln/src/circuit/iden3calc/graph.rs:246-272   graph::evaluate  (23 414 nodes → 5 844 wires per proof)
//   rln/src/circuit/qap.rs:30-98                 CircomReduction::witness_map_from_matrices
//
// Data layout: every per-proof vector lives in a matrix [element][proof] (proof index fastest) so
// that the B proofs of a batch are the coalescing dimension: a warp that processes 32 proofs reads
// 32 consecutive 32-byte field elements (two 128-bit loads per lane).  The NTT twiddles are uniform across the warp.
// (The witness VM kernel lives in k_witness_body.cuh / k_witness.cu: it is compiled with the low-latency multiplier.)
#include <cstdlib>

#include "device_api.hpp"
#include "tma.cuh"

namespace zk {

__device__ __forceinline__ Fr load_canonical_fr(const uint8_t* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    u32 c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    u32 m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = FrCfg::p(i);
    while (Fr::raw_cmp(c, m) >= 0) Fr::raw_sub(c, c, m);
    return Fr::from_canonical(c);
}

// externally calculated witness (generate_zk_proof_with_witness, rln/src/protocol/proof.rs:705-732): wire i of proof j goes to the
// node the graph assigns to that wire, so the QAP and the MSMs read it exactly as if k_witness had produced it
__global__ void k_scatter_wires(CircuitDev c, const uint8_t* __restrict__ wires, Fr* __restrict__ vals, u32 B, u32* __restrict__ err) {
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 j = blockIdx.y;
    if (i == 0) err[j] = 0;
    if (i >= c.n_wires) return;
    st_fp(vals + (size_t)c.signals[i] * B + j, load_canonical_fr(wires + ((size_t)j * c.n_wires + i) * 32));
}
void launch_scatter_wires(const CircuitDev& c, const uint8_t* d_wires, Fr* d_vals, u32 B, u32* d_err, cudaStream_t s) {
    k_scatter_wires<<<dim3((c.n_wires + 127) / 128, B), 128, 0, s>>>(c, d_wires, d_vals, B, d_err);
}

// ------------------------------------------------------------------------------------------- A·w, B·w, c = a∘b
// grid: (ceil(B/128), domain).  Row i < n_constraints: sparse dot products (qap.rs:45-52);
// rows n_constraints .. +n_instance of a copy the public wires (qap.rs:54-58); everything else is zero.
__global__ void __launch_bounds__(128) k_matvec(CircuitDev c, const Fr* __restrict__ vals, Fr* __restrict__ a, Fr* __restrict__ b,
                                                Fr* __restrict__ cc, u32 B) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 row = blockIdx.y;
    if (j >= B) return;
    Fr sa = Fr::zero(), sb = Fr::zero(), sc = Fr::zero();
    if (row < c.n_constraints) {
        for (u32 k = c.a_ptr[row]; k < c.a_ptr[row + 1]; k++) {
            Fr w = ld_fp(vals + (size_t)c.signals[c.a_col[k]] * B + j);
            sa += ldg_fp(c.a_val + k) * w;
        }
        for (u32 k = c.b_ptr[row]; k < c.b_ptr[row + 1]; k++) {
            Fr w = ld_fp(vals + (size_t)c.signals[c.b_col[k]] * B + j);
            sb += ldg_fp(c.b_val + k) * w;
        }
        sc = sa * sb;
    } else if (row < c.n_constraints + c.n_instance) {
        sa = ld_fp(vals + (size_t)c.signals[row - c.n_constraints] * B + j);
    }
    const size_t o = (size_t)row * B + j;
    st_fp(a + o, sa);
    st_fp(b + o, sb);
    st_fp(cc + o, sc);
}

// ------------------------------------------------------------------------------------------- batched radix-2 NTT
// Decimation-in-frequency stage (natural order in → bit-reversed out after all stages):
//   u = x[k+j], v = x[k+j+half];  x[k+j] = u+v;  x[k+j+half] = (u−v)·tw[j·stride]
// grid: (ceil(B/128), n/2).  The twiddle is uniform per block.
__global__ void __launch_bounds__(128) k_ntt_dif_stage(Fr* __restrict__ x, u32 B, u32 half, u32 stride, const Fr* __restrict__ tw) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const u32 t = blockIdx.y;
    const u32 j = t & (half - 1);
    const u32 i0 = ((t - j) << 1) + j;
    Fr* p0 = x + (size_t)i0 * B + p;
    Fr* p1 = p0 + (size_t)half * B;
    Fr u = ld_fp(p0), v = ld_fp(p1);
    st_fp(p0, u + v);
    Fr d = u - v;
    if (j) d = d * ldg_fp(tw + (size_t)j * stride);
    st_fp(p1, d);
}
// Decimation-in-time stage (bit-reversed in → natural out):
//   u = x[k+j], v = x[k+j+half]·tw[j·stride];  x[k+j] = u+v;  x[k+j+half] = u−v
__global__ void __launch_bounds__(128) k_ntt_dit_stage(Fr* __restrict__ x, u32 B, u32 half, u32 stride, const Fr* __restrict__ tw) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const u32 t = blockIdx.y;
    const u32 j = t & (half - 1);
    const u32 i0 = ((t - j) << 1) + j;
    Fr* p0 = x + (size_t)i0 * B + p;
    Fr* p1 = p0 + (size_t)half * B;
    Fr u = ld_fp(p0), v = ld_fp(p1);
    if (j) v = v * ldg_fp(tw + (size_t)j * stride);
    st_fp(p0, u + v);
    st_fp(p1, u - v);
}
// x[pos] *= factor[pos]   (coset shift g^{rev(pos)} and the 1/n of the inverse transform, fused)
__global__ void __launch_bounds__(128) k_scale_rows(Fr* __restrict__ x, u32 B, const Fr* __restrict__ factor) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    Fr* q = x + (size_t)blockIdx.y * B + p;
    st_fp(q, ld_fp(q) * ldg_fp(factor + blockIdx.y));
}
// h = a·b − c  (qap.rs:84,93-95), written over a
__global__ void __launch_bounds__(128) k_h_combine(Fr* __restrict__ a, const Fr* __restrict__ b, const Fr* __restrict__ c, u32 B) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const size_t o = (size_t)blockIdx.y * B + p;
    st_fp(a + o, ld_fp(a + o) * ld_fp(b + o) - ld_fp(c + o));
}

// Two fused decimation-in-frequency stages (halves h and h/2): one thread owns the 4 points
// {i, i+h/2, i+h, i+3h/2} of one proof, so the data makes one HBM round trip per two stages.
__global__ void __launch_bounds__(128) k_ntt_dif_r4(Fr* __restrict__ x, u32 B, u32 h, u32 stride, const Fr* __restrict__ tw) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const u32 q = h >> 1, t = blockIdx.y;
    const u32 j = t & (q - 1);
    const u32 base = ((t - j) << 2) + j;  // (t / q) · 2h + j
    Fr* p0 = x + (size_t)base * B + p;
    Fr* p1 = p0 + (size_t)q * B;
    Fr* p2 = p0 + (size_t)h * B;
    Fr* p3 = p2 + (size_t)q * B;
    Fr x0 = ld_fp(p0), x1 = ld_fp(p1), x2 = ld_fp(p2), x3 = ld_fp(p3);
    // stage A (half = h): (x0,x2) with ω^{j·stride}, (x1,x3) with ω^{(j+q)·stride}
    Fr u0 = x0 + x2, u2 = x0 - x2, u1 = x1 + x3, u3 = x1 - x3;
    if (j) u2 = u2 * ldg_fp(tw + (size_t)j * stride);
    u3 = u3 * ldg_fp(tw + (size_t)(j + q) * stride);
    // stage B (half = q, stride doubled): both pairs with ω^{j·2·stride}
    Fr y0 = u0 + u1, y1 = u0 - u1, y2 = u2 + u3, y3 = u2 - u3;
    if (j) {
        const Fr wb = ldg_fp(tw + (size_t)j * 2 * stride);
        y1 = y1 * wb;
        y3 = y3 * wb;
    }
    st_fp(p0, y0); st_fp(p1, y1); st_fp(p2, y2); st_fp(p3, y3);
}
// Two fused decimation-in-time stages (halves m and 2m): points {i, i+m, i+2m, i+3m}
__global__ void __launch_bounds__(128) k_ntt_dit_r4(Fr* __restrict__ x, u32 B, u32 m, u32 stride, const Fr* __restrict__ tw) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= B) return;
    const u32 t = blockIdx.y;
    const u32 j = t & (m - 1);
    const u32 base = ((t - j) << 2) + j;  // (t / m) · 4m + j
    Fr* p0 = x + (size_t)base * B + p;
    Fr* p1 = p0 + (size_t)m * B;
    Fr* p2 = p1 + (size_t)m * B;
    Fr* p3 = p2 + (size_t)m * B;
    Fr x0 = ld_fp(p0), x1 = ld_fp(p1), x2 = ld_fp(p2), x3 = ld_fp(p3);
    // stage A (half = m, stride): (x0,x1) and (x2,x3) with ω^{j·stride}
    if (j) {
        const Fr wa = ldg_fp(tw + (size_t)j * stride);
        x1 = x1 * wa;
        x3 = x3 * wa;
    }
    Fr a0 = x0 + x1, a1 = x0 - x1, a2 = x2 + x3, a3 = x2 - x3;
    // stage B (half = 2m, stride/2): (a0,a2) with ω^{j·stride/2}, (a1,a3) with ω^{(j+m)·stride/2}
    const u32 sb = stride >> 1;
    if (j) a2 = a2 * ldg_fp(tw + (size_t)j * sb);
    a3 = a3 * ldg_fp(tw + (size_t)(j + m) * sb);
    st_fp(p0, a0 + a2); st_fp(p2, a0 - a2); st_fp(p1, a1 + a3); st_fp(p3, a1 - a3);
}

static void ntt_dif(Fr* x, u32 log_n, u32 B, const Fr* tw, cudaStream_t s) {
    const u32 n = 1u << log_n;
    u32 half = n / 2, stride = 1;
    while (half >= 2) {  // two stages per pass
        k_ntt_dif_r4<<<dim3((B + 127) / 128, n / 4), 128, 0, s>>>(x, B, half, stride, tw);
        half >>= 2;
        stride <<= 2;
    }
    if (half == 1) k_ntt_dif_stage<<<dim3((B + 127) / 128, n / 2), 128, 0, s>>>(x, B, 1, stride, tw);
}
static void ntt_dit(Fr* x, u32 log_n, u32 B, const Fr* tw, cudaStream_t s) {
    const u32 n = 1u << log_n;
    u32 half = 1, stride = n / 2;
    while (half * 2 < n) {
        k_ntt_dit_r4<<<dim3((B + 127) / 128, n / 4), 128, 0, s>>>(x, B, half, stride, tw);
        half <<= 2;
        stride >>= 2;
    }
    if (half < n) k_ntt_dit_stage<<<dim3((B + 127) / 128, n / 2), 128, 0, s>>>(x, B, half, stride, tw);
}
u32 ntt_launches_per_transform(u32 log_n) { return log_n / 2 + (log_n & 1); }

// ---- tiled NTTs: TMA bulk copies into shared memory, 6–7 butterfly stages per HBM round trip ---------------------------------------
// The pass-per-two-stages transforms above make 7 HBM round trips per transform, 15 per buffer with the coset scaling (47 launches
// per QAP), and move 7× the algorithmic 3 MiB per proof.  For small batches — where those launches ARE the cost — the
// iNTT → coset shift → NTT chain of one buffer is THREE kernels, each moving a tile of NTT_P = 16 proofs through shared memory:
//   A  decimation-in-frequency, the outer L − 6 stages: tile = the 2^(L−6) rows {i0 + 64·k} (64 tiles per 16 proofs);
//   B  a contiguous block of 64 rows: the last 6 DIF stages, × g^rev(i)/n, the first 6 decimation-in-time stages;
//   C  decimation-in-time, the outer L − 6 stages, same tiles as A.
// Rows arrive by 1-D TMA bulk copies (cp.async.bulk, 512 B per row and tile, completing on one mbarrier), so do the tile's
// twiddles — stored tile-major on the host so that each tile's are contiguous (64·R + 2·64 values: 0.3 MB) — and results leave by bulk
// stores (cp.async.bulk.global.shared::cta + bulk_group).  Same butterflies, same twiddles, same order: bit-identical output.
constexpr u32 NTT_P = 16;          // proofs per tile
constexpr u32 NTT_TILED_MAX_BATCH = 256;
constexpr u32 NTT_MID = 64;        // rows of the middle block (6 stages)
static __device__ __forceinline__ void tma_store_1d(void* gmem_dst, const void* smem_src, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the writes have landed, not just the shared-memory reads
}
static __device__ __forceinline__ void mbar_expect(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
static __device__ __forceinline__ void tma_load_1d_noexpect(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
static __device__ __forceinline__ Fr sm_ld(const uint4* sm, u32 row, u32 p) {
    const uint4 lo = sm[(row * NTT_P + p) * 2], hi = sm[(row * NTT_P + p) * 2 + 1];
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
static __device__ __forceinline__ void sm_st(uint4* sm, u32 row, u32 p, const Fr& v) {
    sm[(row * NTT_P + p) * 2] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    sm[(row * NTT_P + p) * 2 + 1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
static __device__ __forceinline__ Fr sm_tw(const uint4* tw, u32 i) {
    const uint4 lo = tw[2 * i], hi = tw[2 * i + 1];
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w; r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
// outer stages.  DIT = false: kernel A (stages with local half R/2 … 1, twiddle on the difference); DIT = true: kernel C (local half
// 1 … R/2, twiddle on the second operand).  tw_tiles: [64][R] values, tile i0's at offset i0·R, stage with local half lh at R − 2·lh.
// grid: (64, ceil(B / NTT_P)); dynamic shared memory: R·NTT_P·32 + R·32 + 8 bytes
template <bool DIT>
__global__ void __launch_bounds__(256) k_ntt_outer(Fr* __restrict__ x, u32 B, u32 R, const Fr* __restrict__ tw_tiles) {
    extern __shared__ __align__(128) uint4 sm[];
    uint4* tw = sm + (size_t)R * NTT_P * 2;
    u64* bar = reinterpret_cast<u64*>(tw + (size_t)R * 2);
    const u32 i0 = blockIdx.x, p0 = blockIdx.y * NTT_P;
    const u32 pv = B - p0 < NTT_P ? B - p0 : NTT_P;            // proofs really in this tile
    const u32 row_bytes = pv * 32;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(bar, R * row_bytes + R * 32);
    }
    __syncthreads();
    if (threadIdx.x < R) tma_load_1d_noexpect(sm + (size_t)threadIdx.x * NTT_P * 2, x + (size_t)(i0 + NTT_MID * threadIdx.x) * B + p0, row_bytes, bar);
    if (threadIdx.x == 255) tma_load_1d_noexpect(tw, tw_tiles + (size_t)i0 * R, R * 32, bar);
    mbar_wait(bar, 0);
    const u32 p = threadIdx.x & (NTT_P - 1), q = threadIdx.x / NTT_P;   // 16 butterfly lanes per proof
    const u32 pairs = R / 2;
    for (u32 st = 0; (1u << st) < R; st++) {
        const u32 lh = DIT ? (1u << st) : (pairs >> st);
        if (p < pv)
            for (u32 t = q; t < pairs; t += 256 / NTT_P) {
                const u32 kj = t & (lh - 1), k0 = ((t - kj) << 1) + kj, k1 = k0 + lh;
                const Fr w = sm_tw(tw, R - 2 * lh + kj);
                Fr u = sm_ld(sm, k0, p), v = sm_ld(sm, k1, p);
                if (DIT) {
                    v = v * w;
                    sm_st(sm, k0, p, u + v);
                    sm_st(sm, k1, p, u - v);
                } else {
                    sm_st(sm, k0, p, u + v);
                    sm_st(sm, k1, p, (u - v) * w);
                }
            }
        __syncthreads();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes above → visible to the bulk stores below
    __syncthreads();
    if (threadIdx.x < R) {
        tma_store_1d(x + (size_t)(i0 + NTT_MID * threadIdx.x) * B + p0, sm + (size_t)threadIdx.x * NTT_P * 2, row_bytes);
        tma_store_commit_wait();
    }
}
// middle block: rows [64·blk, 64·blk + 64).  tw_mid: [2][64] values (DIF twiddles then DIT twiddles, stage with half h at 64 − 2·h);
// factor: the per-row coset / 1/n factor in bit-reversed position order (CircuitDev::coset).
// grid: (n / 64, ceil(B / NTT_P)); dynamic shared memory: 64·NTT_P·32 + 3·64·32 + 8 bytes
__global__ void __launch_bounds__(256) k_ntt_middle(Fr* __restrict__ x, u32 B, const Fr* __restrict__ tw_mid, const Fr* __restrict__ factor) {
    extern __shared__ __align__(128) uint4 sm[];
    uint4* tw = sm + (size_t)NTT_MID * NTT_P * 2;       // 128 twiddles
    uint4* fac = tw + 2 * 2 * NTT_MID;                  // 64 factors
    u64* bar = reinterpret_cast<u64*>(fac + 2 * NTT_MID);
    const u32 r0 = blockIdx.x * NTT_MID, p0 = blockIdx.y * NTT_P;
    const u32 pv = B - p0 < NTT_P ? B - p0 : NTT_P;
    const u32 row_bytes = pv * 32;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect(bar, NTT_MID * row_bytes + 3 * NTT_MID * 32);
    }
    __syncthreads();
    if (threadIdx.x < NTT_MID) tma_load_1d_noexpect(sm + (size_t)threadIdx.x * NTT_P * 2, x + (size_t)(r0 + threadIdx.x) * B + p0, row_bytes, bar);
    if (threadIdx.x == 254) tma_load_1d_noexpect(tw, tw_mid, 2 * NTT_MID * 32, bar);
    if (threadIdx.x == 255) tma_load_1d_noexpect(fac, factor + r0, NTT_MID * 32, bar);
    mbar_wait(bar, 0);
    const u32 p = threadIdx.x & (NTT_P - 1), q = threadIdx.x / NTT_P;
    const u32 pairs = NTT_MID / 2;
    for (u32 h = pairs; h >= 1; h >>= 1) {           // decimation in frequency, halves 32 … 1
        if (p < pv)
            for (u32 t = q; t < pairs; t += 256 / NTT_P) {
                const u32 kj = t & (h - 1), k0 = ((t - kj) << 1) + kj, k1 = k0 + h;
                const Fr u = sm_ld(sm, k0, p), v = sm_ld(sm, k1, p);
                sm_st(sm, k0, p, u + v);
                sm_st(sm, k1, p, (u - v) * sm_tw(tw, NTT_MID - 2 * h + kj));
            }
        __syncthreads();
    }
    if (p < pv)                                      // · g^rev(row) / n
        for (u32 k = q; k < NTT_MID; k += 256 / NTT_P) sm_st(sm, k, p, sm_ld(sm, k, p) * sm_tw(fac, k));
    __syncthreads();
    for (u32 h = 1; h <= pairs; h <<= 1) {           // decimation in time, halves 1 … 32
        if (p < pv)
            for (u32 t = q; t < pairs; t += 256 / NTT_P) {
                const u32 kj = t & (h - 1), k0 = ((t - kj) << 1) + kj, k1 = k0 + h;
                const Fr u = sm_ld(sm, k0, p), v = sm_ld(sm, k1, p) * sm_tw(tw, NTT_MID + NTT_MID - 2 * h + kj);
                sm_st(sm, k0, p, u + v);
                sm_st(sm, k1, p, u - v);
            }
        __syncthreads();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < NTT_MID) {
        tma_store_1d(x + (size_t)(r0 + threadIdx.x) * B + p0, sm + (size_t)threadIdx.x * NTT_P * 2, row_bytes);
        tma_store_commit_wait();
    }
}
// iNTT (unscaled, DIF) → × coset factors → NTT (DIT) of one [n][B] buffer in three launches
static void ntt_chain_tiled(Fr* x, const CircuitDev& c, u32 B, cudaStream_t s) {
    const u32 R = c.domain / NTT_MID;                          // rows of an outer tile = 2^(L − 6)
    const size_t smem_outer = (size_t)R * NTT_P * 32 + (size_t)R * 32 + 16, smem_mid = (size_t)NTT_MID * NTT_P * 32 + 3 * NTT_MID * 32 + 16;
    const dim3 grid_outer(NTT_MID, (B + NTT_P - 1) / NTT_P), grid_mid(c.domain / NTT_MID, (B + NTT_P - 1) / NTT_P);
    ZK_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_outer<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_outer));
    ZK_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_outer<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_outer));
    ZK_CUDA_CHECK(cudaFuncSetAttribute(k_ntt_middle, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mid));
    k_ntt_outer<false><<<grid_outer, 256, smem_outer, s>>>(x, B, R, c.tw_tile_dif);
    k_ntt_middle<<<grid_mid, 256, smem_mid, s>>>(x, B, c.tw_mid, c.coset);
    k_ntt_outer<true><<<grid_outer, 256, smem_outer, s>>>(x, B, R, c.tw_tile_dit);
    ZK_CUDA_CHECK(cudaGetLastError());
}
static bool ntt_tiled_ok(const CircuitDev& c) { return c.tw_tile_dif && c.log_domain >= 8 && c.log_domain <= 13; }

// kernels one launch_qap enqueues (bench bookkeeping: gpu_launches)
u32 qap_launch_count(const CircuitDev& c, u32 B) {
    if (B <= NTT_TILED_MAX_BATCH && ntt_tiled_ok(c)) return 2 + 3 * 3;
    return 2 + 3 * (2 * ntt_launches_per_transform(c.log_domain) + 1);
}
void launch_qap(const CircuitDev& c, const Fr* d_vals, Fr* d_a, Fr* d_b, Fr* d_c, u32 B, cudaStream_t s) {
    dim3 grid((B + 127) / 128, c.domain);
    k_matvec<<<grid, 128, 0, s>>>(c, d_vals, d_a, d_b, d_c, B);
    Fr* bufs[3] = {d_a, d_b, d_c};
    // Which form?  Measured on a B200 (profiles/r02g_ntt_msm_affine.txt): the tiled chain wins where launches and HBM round trips
    // dominate — single proof 0.54 → 0.37 ms for the whole QAP (11 launches instead of 47) — is level at batch 256 (1.88 / 1.89 ms)
    // and LOSES at batch 4 096 (24.8 → 26.5 ms): there the two-stage passes already keep the multiplier 86 % busy, and a stage in
    // shared memory pays a barrier plus four 128-bit shared accesses per butterfly.  So: tiled up to NTT_TILED_MAX_BATCH proofs.
    const bool tiled = B <= NTT_TILED_MAX_BATCH && ntt_tiled_ok(c);
    for (Fr* x : bufs) {
        if (tiled) { ntt_chain_tiled(x, c, B, s); continue; }
        ntt_dif(x, c.log_domain, B, c.tw_inv, s);            // ifft (unscaled), output bit-reversed
        k_scale_rows<<<grid, 128, 0, s>>>(x, B, c.coset);    // · g^i / n   (qap.rs:72-79)
        ntt_dit(x, c.log_domain, B, c.tw_fwd, s);            // fft on the coset, natural order
    }
    k_h_combine<<<grid, 128, 0, s>>>(d_a, d_b, d_c, B);
}

// test hook: natural-order forward / inverse transform of [n][B] (inverse leaves out the 1/n factor when tw = ω^{-k})
__global__ void k_bitrev_rows(Fr* __restrict__ x, u32 B, u32 log_n) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 i = blockIdx.y;
    const u32 r = __brev(i) >> (32 - log_n);
    if (p >= B || i >= r) return;
    Fr a = ld_fp(x + (size_t)i * B + p), b = ld_fp(x + (size_t)r * B + p);
    st_fp(x + (size_t)i * B + p, b);
    st_fp(x + (size_t)r * B + p, a);
}
void launch_ntt_test(Fr* d_data, u32 log_n, u32 B, bool inverse, const Fr* tw, cudaStream_t s) {
    (void)inverse;
    ntt_dif(d_data, log_n, B, tw, s);
    dim3 grid((B + 127) / 128, 1u << log_n);
    k_bitrev_rows<<<grid, 128, 0, s>>>(d_data, B, log_n);
}

}  // namespace zk
