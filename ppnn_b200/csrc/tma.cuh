// Thin wrappers over the Blackwell/Hopper asynchronous-copy PTX used by the staged
// kernels: 1-D bulk copies (TMA, `cp.async.bulk`) global -> shared that complete on an
// mbarrier, and the mbarrier primitives of a producer / consumer ring.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ppnn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// PPNN_MBAR_HINT_NS > 0: try_wait with a suspend-time hint (the thread sleeps in hardware until the phase completes
// or the hint elapses, instead of coming back to the issue queue); PPNN_MBAR_SLEEP_NS > 0: nanosleep between polls.
#ifndef PPNN_MBAR_HINT_NS
#define PPNN_MBAR_HINT_NS 0
#endif
#ifndef PPNN_MBAR_SLEEP_NS
#define PPNN_MBAR_SLEEP_NS 0
#endif
__device__ __forceinline__ uint32_t mbar_try_wait_addr(uint32_t a, uint32_t parity) {
    uint32_t done;
#if PPNN_MBAR_HINT_NS > 0
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity), "r"((uint32_t)PPNN_MBAR_HINT_NS) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
#endif
    return done;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    while (!mbar_try_wait_addr(a, parity)) {
#if PPNN_MBAR_SLEEP_NS > 0
        __nanosleep(PPNN_MBAR_SLEEP_NS);
#endif
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// the same on a precomputed shared-window address (hot loops: no generic -> shared conversion per step)
__device__ __forceinline__ void mbar_wait_addr(uint32_t a, uint32_t parity) {
    while (!mbar_try_wait_addr(a, parity)) {
#if PPNN_MBAR_SLEEP_NS > 0
        __nanosleep(PPNN_MBAR_SLEEP_NS);
#endif
    }
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t a) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace ppnn
