// 1-D bulk copies (the TMA engine's linear mode: cp.async.bulk → UBLKCP) and the mbarrier calls that go with them, as inline PTX.
// Used by the witness VM (schedule stream) and the tiled NTT kernels (rows, twiddles; bulk stores).
#pragma once
#include "fp.cuh"

namespace zk {

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

}  // namespace zk
