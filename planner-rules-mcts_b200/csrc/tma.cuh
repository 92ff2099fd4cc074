// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers for staging the
// scenario blob (lane polylines, automaton transition tables, parameters) in shared
// memory.  One elected thread arms the barrier with the byte count and issues the
// copy; every thread then waits on the barrier phase.
#pragma once
#include <stdint.h>

namespace brm {

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// Stage `bytes` (multiple of 16, 16-byte aligned on both sides) from global to shared
// memory with one bulk copy; returns once the data is visible to every thread of the
// CTA.  `bar` must be a shared 8-byte-aligned uint64_t used for nothing else in this
// phase; `parity` is the phase bit of this use (0 for the first, flips each reuse).
__device__ __forceinline__ void stage_blob(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                           uint64_t* bar, uint32_t parity, bool init) {
  if (threadIdx.x == 0) {
    if (init) {
      mbar_init(bar, 1);
      fence_mbar_init();
    }
    mbar_expect_tx(bar, bytes);
    tma_load_1d(smem_dst, gmem_src, bytes, bar);
  }
  if (init) __syncthreads();  // barrier object initialised before anyone polls it
  mbar_wait(bar, parity);
}

}  // namespace brm
