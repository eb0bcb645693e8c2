// Bulk asynchronous copies (TMA, non-tensor form) global -> shared with mbarrier completion, sm_100a.
// Used by the HBM-bound kernels that stream per-episode [n, n] factors: one elected thread requests a whole group of
// matrices as ONE contiguous copy (cp.async.bulk, SASS UBLKCP); the copy engine fills shared memory while the CTA's
// warps keep no loads of their own in flight.  Source, destination and byte count must be multiples of 16.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t bulk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bulk_smem_u32(bar)), "r"(arrivals) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // visible to the copy engine (async proxy)
}

// one arrival + the number of bytes the copies bound to this phase will deliver
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bulk_smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   bulk_smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(bulk_smem_u32(bar))
               : "memory");
}

// spin until the phase with this parity has completed (every thread that reads the data waits itself: the wait is
// what orders its shared-memory reads after the copy engine's writes)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bulk_smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
