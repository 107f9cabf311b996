// The witness-graph VM kernel (rln/src/circuit/iden3calc/graph.rs:246-272 graph::evaluate), as a body that a
// translation unit includes after choosing the multiplier: k_prover.cu does not use it any more; k_witness.cu compiles it with
// ZK_MUL_LOWLAT because the kernel is latency-bound at every batch size (one CTA of four lone warps per SM).
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

}  // namespace zk
