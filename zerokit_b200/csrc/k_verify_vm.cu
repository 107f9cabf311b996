// Groth16 verification of ONE proof by ONE CTA: the interpreter of the pairing VM (verify_vm.hpp — the program is traced and
// scheduled on the host once per verifying key) plus the warp that computes vk_x from the public inputs.
//
// Replaces, like k_verify.cu, rln/src/protocol/proof.rs:856-894 (verify_zk_proof) + Proof::deserialize_compressed (:456-470);
// same result codes.  k_verify runs one proof per thread (36 000 dependent Fq products: 22 ms for a single call); here NW warps
// execute ≈ 1 600 levels of up to 128 independent sums of products each, slots in shared memory, a named barrier between levels.
// A proof this kernel cannot decide (a point at infinity, an exceptional addition) is reported as 3 and re-run by k_verify.
#include "device_api.hpp"
#include "fixed_base.cuh"
#include "tma.cuh"
#include "verify_vm_special.cuh"

namespace zk {
using namespace pvm;

__device__ __forceinline__ Fq ld_slot(const Fq* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 a = q[0], b = q[1];
    Fq r;
    r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w;
    r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
__device__ __forceinline__ void st_slot(Fq* p, const Fq& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// one lane's sum of N products: operands from the slot file, the a operand complemented (q − a) and / or shifted as its term
// word says — integers below 2^256 either way (a ≤ q, shift ≤ 2)
template <int N>
__device__ __forceinline__ Fq vm_dot(const u32* t, const Fq* slots) {
    Fq A[N], B[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        A[k] = ld_slot(slots + (t[k] & 0xfff));
        B[k] = ld_slot(slots + ((t[k] >> 12) & 0xfff));
        if ((t[k] >> 24) & 1) A[k] = A[k].neg_lazy();
        const u32 sh = (t[k] >> 25) & 3;
#pragma unroll
        for (int i = 7; i > 0; i--) A[k].l[i] = __funnelshift_l(A[k].l[i - 1], A[k].l[i], sh);
        A[k].l[0] <<= sh;
    }
    return Fq::dot_wide<N>(A, B);
}
// out of line (one copy of each width; inlined into the level loop the shuffles of the combine step lost their converged fast
// path: 3 400 → 7 000 cycles per level).  The caller must have no global loads in flight into registers at the call: the call waits
// for them (registers the callee may clobber) — which is why the records are staged in shared memory by bulk copies, not prefetched
// into registers (that cost 1 100 cycles of L2 latency per level, ncu round 2).
struct TermWords { u32 t[NMAX]; };
__device__ __noinline__ Fq vm_dot_n(u32 N, TermWords tw, const Fq* slots) {
    const u32* t = tw.t;
    switch (N) {
        case 1: return vm_dot<1>(t, slots);
        case 2: return vm_dot<2>(t, slots);
        case 3: return vm_dot<3>(t, slots);
        case 4: return vm_dot<4>(t, slots);
        case 5: return vm_dot<5>(t, slots);
        case 6: return vm_dot<6>(t, slots);
        case 7: return vm_dot<7>(t, slots);
        default: return vm_dot<8>(t, slots);
    }
}
constexpr u32 PVM_RING = 4;                                   // levels of records in flight
constexpr u32 PVM_REC_BYTES = REC_WORDS * LANES * 4;
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// vk_x = γ_abc[0] + Σ xᵢ·γ_abc[i+1] on one warp: every public input's windows are spread over 32 / n_public lanes (a lane adds
// its share of table entries with complete mixed additions), a shuffle tree of complete additions joins the partial sums
__device__ void vkx_warp(const VerifyKeyDev& vk, const uint8_t* __restrict__ publics, Fq* slots, u32* status) {
    const u32 lane = threadIdx.x & 31;
    const u32 per = vk.n_public ? 32 / vk.n_public : 1;                 // lanes per public input (≥ 1 for n_public ≤ 32)
    const u32 pub = lane / (per ? per : 1), part = lane % (per ? per : 1);
    G1XYZZ acc = G1XYZZ::infinity();
    if (per && pub < vk.n_public) {
        u32 x[8], m[8];
        for (int i = 0; i < 8; i++) {
            const uint8_t* p = publics + pub * 32 + 4 * i;
            x[i] = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
            m[i] = FrCfg::p(i);
        }
        while (Fr::raw_cmp(x, m) >= 0) Fr::raw_sub(x, x, m);
        const int K = vk.gK, c = vk.gc;
        const int chunk = (K + (int)per - 1) / (int)per, w0 = (int)part * chunk, w1 = min(K, w0 + chunk);
        const u32 half = 1u << (c - 1);
        const G1Affine* tb = vk.gamma_tab + (size_t)pub * ((size_t)K << (c - 1));
        u32 carry = 0;
        for (int w = 0; w < w1; w++) {
            const int d = window_digit(x, w, c, carry);   // the carry of the signed digits runs from window 0
            if (w < w0 || d == 0) continue;
            G1Affine pt = ld_point<Fq>(tb + (size_t)w * half + ((d < 0 ? -d : d) - 1));
            if (d < 0) pt.y = pt.y.neg();
            acc.add_affine(pt);
        }
        if (part == per - 1 && carry) {   // a carry out of the last window cannot happen for scalars below r with K·c ≥ 255
            *status = ST_FALLBACK;
        }
    }
    if (lane == 0) acc.add(G1XYZZ::from_affine(ld_point<Fq>(vk.gamma_abc)));
    for (int off = 16; off > 0; off >>= 1) {
        G1XYZZ o;
        for (int i = 0; i < 8; i++) {
            o.X.l[i] = __shfl_down_sync(0xffffffffu, acc.X.l[i], off);
            o.Y.l[i] = __shfl_down_sync(0xffffffffu, acc.Y.l[i], off);
            o.ZZ.l[i] = __shfl_down_sync(0xffffffffu, acc.ZZ.l[i], off);
            o.ZZZ.l[i] = __shfl_down_sync(0xffffffffu, acc.ZZZ.l[i], off);
        }
        if (lane < off) acc.add(o);
    }
    if (lane == 0) {
        if (acc.is_inf()) *status = ST_FALLBACK;   // e(O, γ) = 1: k_verify drops the factor
        st_slot(slots + S_VX, acc.X); st_slot(slots + S_VY, acc.Y); st_slot(slots + S_VZZ, acc.ZZ); st_slot(slots + S_VZZZ, acc.ZZZ);
    }
}

__global__ void __launch_bounds__(32 * (NW + 1)) k_verify_vm(VerifyVmDev vm, VerifyKeyDev vk, const uint8_t* __restrict__ proofs,
                                                             const uint8_t* __restrict__ publics, size_t n, uint8_t* __restrict__ ok) {
    extern __shared__ __align__(16) uint4 smem_raw[];
    Fq* slots = reinterpret_cast<Fq*>(smem_raw);
    __shared__ u32 s_status;
    __shared__ ProofFlags s_flags;
    const size_t j = blockIdx.x;
    const u32 tid = threadIdx.x, warp = tid >> 5;
    {   // constants (the pinned slots, the per-key line coefficients): one coalesced copy
        const uint4* src = reinterpret_cast<const uint4*>(vm.consts);
        for (u32 i = tid; i < vm.n_const * 2; i += blockDim.x) smem_raw[i] = src[i];
    }
    if (tid == 0) s_status = ST_RUNNING;
    __syncthreads();
    if (tid == 0) {
        ProofFlags fl{0};
        s_status = pv_prologue(proofs + 128 * j, slots, fl);
        s_flags = fl;
    }
    __syncthreads();
    if (s_status != ST_RUNNING) {
        if (tid == 0) ok[j] = (uint8_t)(s_status == ST_INVALID ? 0 : s_status);
        return;
    }
    if (warp == NW) {
        vkx_warp(vk, publics + j * vk.n_public * 32, slots, &s_status);
        __threadfence_block();
        bar_sync(2, 32 * (NW + 1));
        return;
    }
    // the program streams through a ring of PVM_RING levels in shared memory: one elected thread keeps the bulk copies going
    // (cp.async.bulk + mbarrier, as the witness VM does for its schedule), every lane reads its own words of the current level
    u32* ring = reinterpret_cast<u32*>(slots + vm.n_slots);
    u64* full = reinterpret_cast<u64*>(ring + PVM_RING * REC_WORDS * LANES);
    if (tid == 0) {
        for (u32 s = 0; s < PVM_RING; s++) mbar_init(full + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (u32 s = 0; s < PVM_RING && s < vm.n_levels; s++)
            tma_load_1d(ring + s * REC_WORDS * LANES, vm.code + (size_t)s * REC_WORDS * LANES, PVM_REC_BYTES, full + s);
    }
    bar_sync(1, LANES);
    for (u32 l = 0; l < vm.n_levels; l++) {
        const u32 st = l % PVM_RING;
        mbar_wait(full + st, (l / PVM_RING) & 1);
        u32 w[REC_WORDS];
#pragma unroll
        for (int k = 0; k < REC_WORDS; k++) w[k] = ring[st * REC_WORDS * LANES + k * LANES + tid];
        if (vm.trace && tid == 0) vm.trace[l] = clock64();
        const u32 special = (w[1] >> 8) & 0xff;
        if (special) {
            if (special == SP_VKX) {
                bar_sync(2, 32 * (NW + 1));
            } else if (special == SP_EXP) {
                if (w[0] & W0_STORE) {   // lanes 0..EXP_LANES−1 of warp 0: each raises its own value, table in its scratch slots
                    const Fq r = pv_pow(ld_slot(slots + (w[2] & 0xfff)), vm.exps[(w[1] >> 16) & 3], slots + ((w[2] >> 12) & 0xfff));
                    st_slot(slots + (w[0] & 0xfff), r);
                }
            } else if (warp == 0) {
                u32 args[32];
#pragma unroll
                for (int k = 0; k < 32; k++) args[k] = __shfl_sync(0xffffffffu, w[0], k);
                if (tid == 0) {
                    if (special == SP_SELECT) {
                        const u32 st = pv_select(slots, args, s_flags);
                        if (st != ST_RUNNING) s_status = st;
                    } else {
                        s_status = pv_final(slots, args);
                    }
                }
            }
            bar_sync(1, LANES);
            if (s_status != ST_RUNNING) break;   // uniform: written before the barrier
        } else {
            const u32 N = w[1] & 15;
            if (N) {
                TermWords tw;
#pragma unroll
                for (int k = 0; k < NMAX; k++) tw.t[k] = w[2 + k];
                Fq r = vm_dot_n(N, tw, slots);
                const u32 nsub = (w[1] >> 4) & 3;
                for (u32 s = 1; s < nsub; s++) Fq::cond_sub_p(r.l);
                if (w[1] & 64) {
                    Fq o;
#pragma unroll
                    for (int i = 0; i < 8; i++) o.l[i] = __shfl_xor_sync(0xffffffffu, r.l[i], 16);
                    if (w[0] & W0_COMBINE) r = r + o;
                    if (w[1] & 128) {
#pragma unroll
                        for (int i = 0; i < 8; i++) o.l[i] = __shfl_xor_sync(0xffffffffu, r.l[i], 8);
                        if (w[0] & W0_COMBINE4) r = r + o;
                    }
                }
                if (w[0] & W0_STORE) st_slot(slots + (w[0] & 0xfff), r);
            }
            bar_sync(1, LANES);
        }
        // every lane is past this level's records: their buffer takes the level PVM_RING ahead
        if (tid == 0 && l + PVM_RING < vm.n_levels)
            tma_load_1d(ring + st * REC_WORDS * LANES, vm.code + (size_t)(l + PVM_RING) * REC_WORDS * LANES, PVM_REC_BYTES, full + st);
    }
    if (vm.trace && tid == 0) vm.trace[vm.n_levels] = clock64();
    if (tid == 0) {
        const u32 st = s_status;
        ok[j] = (uint8_t)(st == ST_VALID ? 1 : st == ST_INVALID ? 0 : st == ST_RUNNING ? ST_FALLBACK : st);
    }
}

void launch_verify_vm(const VerifyVmDev& prog, const VerifyKeyDev& vk, const uint8_t* d_proofs, const uint8_t* d_publics, size_t n, uint8_t* d_ok,
                      cudaStream_t s) {
    if (!n) return;
    const size_t smem = (size_t)prog.n_slots * sizeof(Fq) + (size_t)PVM_RING * PVM_REC_BYTES + PVM_RING * sizeof(u64);
    ZK_CUDA_CHECK(cudaFuncSetAttribute(k_verify_vm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_verify_vm<<<(unsigned)n, 32 * (NW + 1), smem, s>>>(prog, vk, d_proofs, d_publics, n, d_ok);
    ZK_CUDA_CHECK(cudaGetLastError());
}

}  // namespace zk
