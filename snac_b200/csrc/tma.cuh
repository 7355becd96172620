// TMA (cp.async.bulk.tensor) + mbarrier helpers shared by the z-marching stencil kernels (mom_tma.cu, pois_tma.cu).
//
// Addressing convention of the tensor maps (tma_make_map): a map describes one ghosted field as a 3-D tensor of
// doubles with the row pitch of Grid3; its base must be 16-byte aligned and so must the first element of every box
// row (a box that starts 8 bytes off raises "illegal instruction" at the UTMALDG; tools/probe/tma_probe3.cu).  Grid3
// aligns the first INTERIOR element (1,j,k) of a row on 128 bytes, so two bases are used:
//   ghosted  maps: base = element (-1,0,0)  -> element (i,j,k) has the coordinates (i+1, j, k); boxes start at odd i
//   interior maps: base = element ( 1,0,0)  -> element (i,j,k) has the coordinates (i-1, j, k); boxes start at odd i
// Coordinates outside the tensor (negative, or beyond the extents) are legal: the box is filled with zeros there.
#pragma once
#include <cuda.h>

#include "common.h"

namespace snac {
namespace tma {

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra MBAR_DONE;\n"
      "bra MBAR_WAIT;\n"
      "MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// after mbarrier.init by one thread, before the first TMA that signals the barriers
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace tma

// Tensor map of the field whose (0,0,0) corner is `corner`, for boxes of bx x by x 1 doubles; interior: see above.
// false when the driver entry point is missing or the encoding is refused (the callers fall back to their
// one-thread-per-cell kernels).
bool tma_make_map(void *out128, const double *corner, const Grid3 &g, int bx, int by, bool interior);

}  // namespace snac
