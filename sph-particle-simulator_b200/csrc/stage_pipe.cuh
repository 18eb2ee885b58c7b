// Producer/consumer plumbing of the staged pair kernels (pair_stage.cu): mbarrier rings and 1-D bulk asynchronous
// copies (cp.async.bulk, the TMA engine's non-tensor form) from global into shared memory, sm_90+ PTX.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sphb {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
// upper bound of the time the hardware may keep a waiting warp suspended before try_wait returns false: without it the
// default limit is short and the retry loop burns issue slots (measured: 19 % of the executed instructions)
constexpr uint32_t kSuspendHintNs = 10000000u;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0u;
}
// blocks until the phase of the given parity has completed (a fresh barrier counts as having completed parity 1)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// bytes: a multiple of 16; src and dst 16-byte aligned.  Completion is signalled on `bar` as `bytes` of transaction count.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

}  // namespace

}  // namespace sphb
