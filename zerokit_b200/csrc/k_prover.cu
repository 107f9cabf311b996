// Batched witness generation and QAP witness-map kernels (sm_100a).
//
// Path covered:
//   rln/src/circuit/iden3calc/graph.rs:246-272   graph::evaluate  (23 414 nodes → 5 844 wires per proof)
//   rln/src/circuit/qap.rs:30-98                 CircomReduction::witness_map_from_matrices
//
// Data layout: every per-proof vector lives in a matrix [element][proof] (proof index fastest) so
// that the B proofs of a batch are the coalescing dimension: a warp that processes 32 proofs reads
// 32 consecutive 32-byte field elements (two 128-bit loads per lane).  The instruction stream of
// the witness VM and the NTT twiddles are uniform across the warp.
#include <cstdlib>

#include "device_api.hpp"

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

// ------------------------------------------------------------------------------------------- witness VM
// The graph is a 23 414-node program whose longest dependency chain is 10 000 nodes, and a single warp evaluating it runs at
// 0.2 IPC (dependent issue) while three of the four schedulers of its SM idle.  So 32 proofs share a CTA of four warps: the host
// list-schedules the nodes into bundles of ≤ 4 mutually independent nodes (operands in earlier bundles only), warp w evaluates
// slot w of every bundle for its 32 proofs, and a barrier separates bundles.  A value is written to vals[node][B] (the QAP and
// the MSMs read it there) and to a shared-memory ring of the last VM_RING bundles; 77 % of all operands were produced less than
// 16 bundles earlier and the other 23 % are constants, so the critical path never waits for L2.
template <u32 STRIDE = 32>
__device__ __forceinline__ Fr vm_operand(u32 enc, const uint4* ring, const Fr* __restrict__ consts, const Fr* vals, u32 B, u32 j, u32 lane) {
    const u32 src = enc >> 30, idx = enc & 0x3fffffffu;
    if (src == VM_SRC_RING) {
        const uint4 lo = ring[(idx * 2) * STRIDE + lane], hi = ring[(idx * 2 + 1) * STRIDE + lane];
        Fr r;
        r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
        r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
        return r;
    }
    if (src == VM_SRC_CONST) return ldg_fp(consts + idx);
    return ld_fp(vals + (size_t)idx * B + j);
}
// ---- TMA staging of the schedule ---------------------------------------------------------------------------------------------
// The schedule (one 32-byte record per slot and bundle, 128 B per bundle, 1.3 MB for the depth-20 graph) is the kernel's
// instruction stream.  An add-only bundle lasts ≈ 200 cycles, less than an L2 round trip, so a record fetched on demand sets the
// pace of 5 364 of the 10 337 bundles (round 1: 9.94 ms; fetched 16 bundles ahead through registers + shuffles: 8.58 ms).  Here
// one elected thread streams the schedule into shared memory with 1-D bulk copies (cp.async.bulk → UBLKCP, the TMA engine's
// linear mode): VM_STAGES blocks of VM_STAGE_BUNDLES bundles are in flight, each completes on its own mbarrier
// (mbarrier.arrive.expect_tx / complete_tx), and every warp reads its record with two broadcast 128-bit shared loads.  The
// bundle barrier that the dependency chain needs anyway also tells the producer when a block has been consumed.
constexpr u32 VM_STAGE_BUNDLES = 32;                                         // 32 bundles × 4 slots × 32 B = 4 KB per bulk copy
constexpr u32 VM_STAGES = 4;
constexpr u32 VM_STAGE_BYTES = VM_STAGE_BUNDLES * VM_SLOTS * 32;
constexpr size_t VM_RING_BYTES = (size_t)VM_RING * VM_SLOTS * 2 * 32 * sizeof(uint4);   // 64 KB
constexpr size_t VM_SMEM_BYTES = VM_RING_BYTES + (size_t)VM_STAGES * VM_STAGE_BYTES + VM_STAGES * sizeof(u64);

static __device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
static __device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    u32 done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(smem_addr(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// arm the barrier with the byte count and start the bulk copy that will complete it
static __device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, u32 bytes, u64* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)), "l"(gmem_src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// A CTA of four warps carries 32 proofs; warp w evaluates slot w, lane = proof; a CTA barrier separates bundles.  (Round-2
// experiment, measured and removed: ONE warp carrying 8 proofs with lane = slot + 4·proof and __syncwarp instead of the CTA
// barrier was SLOWER — single proof 6.83 → 7.63 ms, batch 4 096 8.33 → 15.2 ms: the four slots of a bundle hold different
// operations and a warp runs divergent lanes one after the other, while four warps run them side by side on four schedulers.)
__global__ void __launch_bounds__(128) k_witness(CircuitDev c, const uint8_t* __restrict__ inputs, Fr* vals, u32 B, u32* __restrict__ err) {
    constexpr u32 PROOFS = 32;                                                              // proofs per CTA = ring stride
    constexpr size_t RING_U4 = (size_t)VM_RING * VM_SLOTS * 2 * PROOFS;
    extern __shared__ __align__(128) uint4 ring[];   // [VM_RING · VM_SLOTS][2][PROOFS]: the two 16-byte halves of a value, proof-contiguous
    uint4* stage = ring + RING_U4;                                                          // [VM_STAGES][VM_STAGE_BUNDLES][VM_SLOTS][2]
    u64* full = reinterpret_cast<u64*>(stage + (size_t)VM_STAGES * VM_STAGE_BYTES / sizeof(uint4));
    const u32 lane = threadIdx.x & 31;
    const u32 slot = threadIdx.x >> 5;
    const u32 pl = lane;                                                                    // proof within the CTA
    const u32 j = blockIdx.x * PROOFS + pl;
    const bool live = j < B;
    const uint8_t* in = inputs + (size_t)(live ? j : 0) * c.n_slots * 32;
    const u32 n_blocks = c.n_bundles / VM_STAGE_BUNDLES;                                   // the host pads the schedule to whole blocks
    const uint8_t* sched = reinterpret_cast<const uint8_t*>(c.sched);
    auto bundle_sync = [] { __syncthreads(); };
    u32 bad = 0;
    if (threadIdx.x == 0) {
        for (u32 s = 0; s < VM_STAGES; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (u32 s = 0; s < VM_STAGES && s < n_blocks; s++)
            tma_load_1d(stage + (size_t)s * VM_STAGE_BYTES / sizeof(uint4), sched + (size_t)s * VM_STAGE_BYTES, VM_STAGE_BYTES, full + s);
    }
    bundle_sync();
    for (u32 blk = 0; blk < n_blocks; blk++) {
        const u32 st = blk % VM_STAGES;
        mbar_wait(full + st, (blk / VM_STAGES) & 1);
        const uint4* recs = stage + (size_t)st * VM_STAGE_BYTES / sizeof(uint4);
#pragma unroll 1
        for (u32 i = 0; i < VM_STAGE_BUNDLES; i++) {
            const u32 b = blk * VM_STAGE_BUNDLES + i;
            const uint4 w0 = recs[2 * (i * VM_SLOTS + slot)];       // kind_op, out, a, b   (one address per slot: broadcast loads)
            if (w0.x != 0xffffffffu && live) {
                const u32 kind = w0.x & 0xff, op = w0.x >> 8;
                Fr v;
                if (kind == VM_DUO) {
                    const Fr x = vm_operand<PROOFS>(w0.z, ring, c.consts, vals, B, j, pl), y = vm_operand<PROOFS>(w0.w, ring, c.consts, vals, B, j, pl);
                    if (op == OP_MUL) v = x * y;
                    else if (op == OP_ADD) v = x + y;
                    else if (op == OP_SUB) v = x - y;
                    else if (!vm_eval_duo(op, x, y, v)) { bad = 1; v = Fr::zero(); }
                } else if (kind == VM_CONST) {
                    v = ldg_fp(c.consts + w0.z);
                } else if (kind == VM_INPUT) {
                    v = load_canonical_fr(in + 32 * w0.z);
                } else if (kind == VM_UNO) {
                    if (op == 0) v = vm_operand<PROOFS>(w0.z, ring, c.consts, vals, B, j, pl).neg();
                    else { bad = 1; v = Fr::zero(); }  // "uno operator Id not implemented" (graph.rs:189-193)
                } else {  // TernCond (graph.rs:216-222)
                    const u32 third = recs[2 * (i * VM_SLOTS + slot) + 1].x;
                    const Fr t = vm_operand<PROOFS>(w0.z, ring, c.consts, vals, B, j, pl);
                    v = t.is_zero() ? vm_operand<PROOFS>(third, ring, c.consts, vals, B, j, pl) : vm_operand<PROOFS>(w0.w, ring, c.consts, vals, B, j, pl);
                }
                const u32 ri = (b % VM_RING) * VM_SLOTS + slot;
                ring[(ri * 2) * PROOFS + pl] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
                ring[(ri * 2 + 1) * PROOFS + pl] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
                st_fp(vals + (size_t)w0.y * B + j, v);
            }
            bundle_sync();
        }
        // every thread is past the last record of this block: its buffer takes the block VM_STAGES ahead
        if (threadIdx.x == 0 && blk + VM_STAGES < n_blocks)
            tma_load_1d(stage + (size_t)st * VM_STAGE_BYTES / sizeof(uint4), sched + (size_t)(blk + VM_STAGES) * VM_STAGE_BYTES, VM_STAGE_BYTES, full + st);
    }
    if (live && bad) atomicOr(err + j, 1u);
}
void launch_witness(const CircuitDev& c, const uint8_t* d_inputs, Fr* d_vals, u32 B, u32* d_err, cudaStream_t s) {
    ZK_CUDA_CHECK(cudaMemsetAsync(d_err, 0, 4 * (size_t)B, s));
    // the shared-memory attribute is per device (a process may drive several GPUs): set on every launch, it is a cheap call
    ZK_CUDA_CHECK(cudaFuncSetAttribute(k_witness, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VM_SMEM_BYTES));
    k_witness<<<(B + 31) / 32, 128, VM_SMEM_BYTES, s>>>(c, d_inputs, d_vals, B, d_err);
    ZK_CUDA_CHECK(cudaGetLastError());
}
u32 vm_schedule_block_bundles() { return VM_STAGE_BUNDLES; }
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
