// tma.cuh — thin inline-PTX wrappers for the Blackwell/Hopper Tensor Memory Accelerator (cp.async.bulk.tensor) and the
// mbarrier that signals its completion.  Used to stage image tiles (u8, 3-D tensor = x, y, frame) into shared memory
// with ONE instruction issued by one thread instead of a per-thread load loop; out-of-image parts of a box arrive as
// zeros and are patched by the border tiles (reflection) where the algorithm needs it.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ivg {

struct TmaMaps { CUtensorMap m[12]; };   // one descriptor per pyramid level (MAX_LEVELS)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "IVG_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra IVG_DONE;\n"
      "bra IVG_WAIT;\n"
      "IVG_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}

// box of the 3-D tensor (x, y, frame) -> shared memory; completion is signalled on `bar` with the byte count
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}

}  // namespace ivg
