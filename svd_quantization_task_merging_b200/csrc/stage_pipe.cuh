// Shared-memory staging pipeline for the streaming kernels: a producer warp moves 1-D chunks of the
// input tensors global -> shared with the TMA bulk-copy engine (cp.async.bulk + mbarrier
// complete_tx), consumer warps wait on "full" barriers and hand slots back through "empty" barriers.
// Keeps STAGES x stage_bytes (up to ~180 KB) of reads in flight per SM without holding them in
// registers.
#pragma once
#include <stdint.h>

namespace svdq {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "SVDQ_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra SVDQ_DONE;\n\t"
        "bra SVDQ_WAIT;\n\t"
        "SVDQ_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ring position of the k-th use of the pipeline
struct PipeState {
    uint32_t stage = 0, phase = 0;
    template <int STAGES> __device__ __forceinline__ void advance() {
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
    }
    __device__ __forceinline__ void advance_n(uint32_t stages) {       // ring depth chosen at launch
        if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
};

}  // namespace svdq
