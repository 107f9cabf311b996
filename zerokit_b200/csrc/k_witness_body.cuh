// The witness-graph VM kernel (rln/src/circuit/iden3calc/graph.rs:246-272 graph::evaluate), as a body that a
// header so that the kernel has a translation unit of its own (k_witness.cu): it is latency-bound at every batch size (one CTA of
// four lone warps per SM) and is tuned separately from the throughput kernels of k_prover.cu.
#pragma once
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

// ------------------------------------------------------------------------------------------- witness VM
// The graph is a 23 414-node program whose longest dependency chain is 10 000 nodes, and a single warp evaluating it runs at
// 0.2 IPC (dependent issue) while three of the four schedulers of its SM idle.  So 32 proofs share a CTA of four warps: the host
// list-schedules the nodes into bundles of ≤ 4 mutually independent nodes (operands in earlier bundles only), warp w evaluates
// slot w of every bundle for its 32 proofs, and a barrier separates bundles.  A value is written to a shared-memory ring of the
// last VM_RING bundles (77 % of all operands were produced less than 16 bundles earlier, the other 23 % are constants) and — only
// if a wire or a far consumer needs it — to vals[node][B], where the QAP and the MSMs read it.
//
// What a bundle costs is the instruction count of its slowest slot (ncu source page, round 2: a lone warp issues one dependent
// instruction every ≈ 5 cycles; a product is ≈ 190 of them, an addition ≈ 40, and the bookkeeping around either was ≈ 80), so the
// bookkeeping is kept off that path: the constant table sits in shared memory next to the ring (one bulk copy at kernel start),
// ring and constant operands are fetched through ONE branch-free address computation, the global store is skipped for the 70 % of
// nodes nobody reads from HBM, and the rare operators go through an out-of-line call that takes its operands by value (passing
// them by reference made ptxas spill both operands to local memory for EVERY node: 4 × STL.128, 8.5 % of the kernel's samples).
constexpr u32 VM_CONST_MAX = 1536;   // constants kept in shared memory (48 KB); a larger table makes the host schedule them as nodes

// operand fetch: ring slot (lane-strided halves), constant (same address for every lane) or — rare — vals[node][B]
__device__ __forceinline__ Fr vm_operand(u32 enc, const uint4* sm, u32 const_base, const Fr* vals, u32 B, u32 j, u32 lane) {
    const u32 src = enc >> 30, idx = enc & 0x3fffffffu;
    if (src == VM_SRC_GLOBAL) return ld_fp(vals + (size_t)idx * B + j);
    const bool is_c = src == VM_SRC_CONST;
    const u32 lo_i = is_c ? const_base + idx * 2 : idx * 64 + lane;
    const uint4 lo = sm[lo_i], hi = sm[lo_i + (is_c ? 1u : 32u)];
    Fr r;
    r.l[0] = lo.x; r.l[1] = lo.y; r.l[2] = lo.z; r.l[3] = lo.w;
    r.l[4] = hi.x; r.l[5] = hi.y; r.l[6] = hi.z; r.l[7] = hi.w;
    return r;
}
// everything but Mul / Add / Sub: out of line, operands by value (registers / param space of the call, not the caller's frame)
__device__ __noinline__ u32 vm_eval_rare(u32 op, Fr x, Fr y, Fr* out) {
    Fr v;
    const bool ok = vm_eval_duo(op, x, y, v);
    *out = ok ? v : Fr::zero();
    return ok ? 0u : 1u;
}
// ---- TMA staging of the schedule ---------------------------------------------------------------------------------------------
// The schedule (one 32-byte record per slot and bundle, 128 B per bundle, 1.3 MB for the depth-20 graph) is the kernel's
// instruction stream.  An add-only bundle lasts a few hundred cycles, less than an L2 round trip, so a record fetched on demand
// sets the pace of half of the bundles (round 1: 9.94 ms; fetched 16 bundles ahead through registers + shuffles: 8.58 ms).  Here
// one elected thread streams the schedule into shared memory with 1-D bulk copies (cp.async.bulk → UBLKCP, the TMA engine's
// linear mode): VM_STAGES blocks of VM_STAGE_BUNDLES bundles are in flight, each completes on its own mbarrier
// (mbarrier.arrive.expect_tx / complete_tx), and every warp reads its record with two broadcast 128-bit shared loads.  The
// bundle barrier that the dependency chain needs anyway also tells the producer when a block has been consumed.
constexpr u32 VM_STAGE_BUNDLES = 32;                                         // 32 bundles × 4 slots × 32 B = 4 KB per bulk copy
constexpr u32 VM_STAGES = 4;
constexpr u32 VM_STAGE_BYTES = VM_STAGE_BUNDLES * VM_SLOTS * 32;
constexpr size_t VM_RING_BYTES = (size_t)VM_RING * VM_SLOTS * 2 * 32 * sizeof(uint4);   // 64 KB
static size_t vm_smem_bytes(u32 n_consts_smem) {
    return VM_RING_BYTES + (size_t)n_consts_smem * 32 + (size_t)VM_STAGES * VM_STAGE_BYTES + (VM_STAGES + 1) * sizeof(u64);
}

// A CTA of four warps carries 32 proofs; warp w evaluates slot w, lane = proof; a CTA barrier separates bundles.  (Round-2
// experiment, measured and removed: ONE warp carrying 8 proofs with lane = slot + 4·proof and __syncwarp instead of the CTA
// barrier was SLOWER — single proof 6.83 → 7.63 ms, batch 4 096 8.33 → 15.2 ms: the four slots of a bundle hold different
// operations and a warp runs divergent lanes one after the other, while four warps run them side by side on four schedulers.)
__global__ void __launch_bounds__(32 * VM_SLOTS) k_witness(CircuitDev c, const uint8_t* __restrict__ inputs, Fr* vals, u32 B, u32* __restrict__ err) {
    extern __shared__ __align__(128) uint4 sm[];   // ring [VM_RING · VM_SLOTS][2][32 lanes] | constants [n][2] | schedule stages | mbarriers
    const u32 const_base = (u32)(VM_RING_BYTES / sizeof(uint4));
    uint4* stage = sm + const_base + (size_t)c.n_consts_smem * 2;                           // [VM_STAGES][VM_STAGE_BUNDLES][VM_SLOTS][2]
    u64* full = reinterpret_cast<u64*>(stage + (size_t)VM_STAGES * VM_STAGE_BYTES / sizeof(uint4));
    u64* cbar = full + VM_STAGES;
    const u32 lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
    const u32 j = blockIdx.x * 32 + lane;
    const bool live = j < B;
    const uint8_t* in = inputs + (size_t)(live ? j : 0) * c.n_slots * 32;
    const u32 n_blocks = c.n_bundles / VM_STAGE_BUNDLES;                                   // the host pads the schedule to whole blocks
    const uint8_t* sched = reinterpret_cast<const uint8_t*>(c.sched);
    u32 bad = 0;
    if (threadIdx.x == 0) {
        for (u32 s = 0; s < VM_STAGES; s++) mbar_init(full + s, 1);
        mbar_init(cbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (c.n_consts_smem) tma_load_1d(sm + const_base, c.consts, c.n_consts_smem * 32, cbar);
        for (u32 s = 0; s < VM_STAGES && s < n_blocks; s++)
            tma_load_1d(stage + (size_t)s * VM_STAGE_BYTES / sizeof(uint4), sched + (size_t)s * VM_STAGE_BYTES, VM_STAGE_BYTES, full + s);
    }
    __syncthreads();
    if (c.n_consts_smem) mbar_wait(cbar, 0);
    for (u32 blk = 0; blk < n_blocks; blk++) {
        const u32 st = blk % VM_STAGES;
        mbar_wait(full + st, (blk / VM_STAGES) & 1);
        const uint4* recs = stage + (size_t)st * VM_STAGE_BYTES / sizeof(uint4);
#pragma unroll 1
        for (u32 i = 0; i < VM_STAGE_BUNDLES; i++) {
            const u32 b = blk * VM_STAGE_BUNDLES + i;
            const uint4 w0 = recs[2 * (i * VM_SLOTS + slot)];       // kind_op, out | store flag, a, b   (warp-uniform address: one broadcast load)
            if (w0.x != 0xffffffffu && live) {
                const u32 kind = w0.x & 0xff, op = w0.x >> 8;
                Fr v;
                if (kind == VM_DUO) {
                    const Fr x = vm_operand(w0.z, sm, const_base, vals, B, j, lane), y = vm_operand(w0.w, sm, const_base, vals, B, j, lane);
                    if (op == OP_MUL) v = w0.z == w0.w ? x.sqr() : x * y;   // x·x (two of the three products of every x⁵ S-box): the dedicated squaring, 682 instead of 865 cycles
                    else if (op == OP_ADD) v = x + y;
                    else if (op == OP_SUB) v = x - y;
                    else bad |= vm_eval_rare(op, x, y, &v);
                } else if (kind == VM_CONST) {
                    v = ldg_fp(c.consts + w0.z);
                } else if (kind == VM_INPUT) {
                    v = load_canonical_fr(in + 32 * w0.z);
                } else if (kind == VM_UNO) {
                    if (op == 0) v = vm_operand(w0.z, sm, const_base, vals, B, j, lane).neg();
                    else { bad = 1; v = Fr::zero(); }  // "uno operator Id not implemented" (graph.rs:189-193)
                } else if (op == VM_TRES_FMA) {   // a·b + c (the rewrite's fused node: the sum rides in the product's bundle)
                    const u32 third = recs[2 * (i * VM_SLOTS + slot) + 1].x;
                    const Fr z = vm_operand(third, sm, const_base, vals, B, j, lane);   // fetched before the product, not after it
                    const Fr x = vm_operand(w0.z, sm, const_base, vals, B, j, lane), y = vm_operand(w0.w, sm, const_base, vals, B, j, lane);
                    v = (w0.z == w0.w ? x.sqr() : x * y) + z;
                } else {  // TernCond (graph.rs:216-222)
                    const u32 third = recs[2 * (i * VM_SLOTS + slot) + 1].x;
                    const Fr t = vm_operand(w0.z, sm, const_base, vals, B, j, lane);
                    v = t.is_zero() ? vm_operand(third, sm, const_base, vals, B, j, lane) : vm_operand(w0.w, sm, const_base, vals, B, j, lane);
                }
                const u32 ri = (b % VM_RING) * VM_SLOTS + slot;
                sm[(ri * 2) * 32 + lane] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
                sm[(ri * 2 + 1) * 32 + lane] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
                if (w0.y >> 31) st_fp(vals + (size_t)(w0.y & 0x7fffffffu) * B + j, v);   // a wire, or read from HBM by a far consumer
            }
            __syncthreads();
        }
        // every warp is past the last record of this block: its buffer takes the block VM_STAGES ahead
        if (threadIdx.x == 0 && blk + VM_STAGES < n_blocks)
            tma_load_1d(stage + (size_t)st * VM_STAGE_BYTES / sizeof(uint4), sched + (size_t)(blk + VM_STAGES) * VM_STAGE_BYTES, VM_STAGE_BYTES, full + st);
    }
    if (live && bad) atomicOr(err + j, 1u);
}
void launch_witness(const CircuitDev& c, const uint8_t* d_inputs, Fr* d_vals, u32 B, u32* d_err, cudaStream_t s) {
    ZK_CUDA_CHECK(cudaMemsetAsync(d_err, 0, 4 * (size_t)B, s));
    // the shared-memory attribute is per device (a process may drive several GPUs): set on every launch, it is a cheap call
    const size_t smem = vm_smem_bytes(c.n_consts_smem);
    ZK_CUDA_CHECK(cudaFuncSetAttribute(k_witness, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_witness<<<(B + 31) / 32, 32 * VM_SLOTS, smem, s>>>(c, d_inputs, d_vals, B, d_err);
    ZK_CUDA_CHECK(cudaGetLastError());
}
u32 vm_schedule_block_bundles() { return VM_STAGE_BUNDLES; }
u32 vm_const_smem_max() { return VM_CONST_MAX; }

}  // namespace zk
