/* nnsp_tma.cuh -- TMA bulk copy (cp.async.bulk) of a weight image into shared memory. */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nnsp {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* Stage `bytes` (multiple of 16) from global to shared with one TMA bulk copy tracked by an
 * mbarrier; every thread of the CTA waits on the barrier's phase 0. */
__device__ __forceinline__ void tma_stage_weights(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
        uint32_t off = 0;
        while (off < bytes) {
            const uint32_t n = (bytes - off) > 32768u ? 32768u : (bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32((char *)dst + off)), "l"((const char *)src + off), "r"(n), "r"(smem_u32(bar)) : "memory");
            off += n;
        }
    }
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    }
}


}  // namespace nnsp
