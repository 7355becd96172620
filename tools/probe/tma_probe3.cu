// morphing the working sample (tma_probe2.cu) towards the raw-PTX form, one step per -DSTEP
#include <cuda.h>
#include <cuda/barrier>
#include <cuda_runtime.h>
#include <cstdio>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int SW = 32, SH = 8;
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int *out) {
#if STEP >= 4
  extern __shared__ __align__(128) int dyn[];
  int (*smem_buffer)[SW] = reinterpret_cast<int (*)[SW]>(dyn);
#else
  __shared__ alignas(128) int smem_buffer[SH][SW];
#endif
#if STEP == 0
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else token = bar.arrive();
  bar.wait(std::move(token));
#else
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#if STEP >= 2
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(SW * SH * 4) : "memory");
#endif
#if STEP >= 3
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(&smem_buffer[0][0])), "l"(&tensor_map), "r"(smem_u32(&bar)), "r"(x), "r"(y) : "memory");
#else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(&smem_buffer[0][0])), "l"(&tensor_map), "r"(x), "r"(y), "r"(smem_u32(&bar)) : "memory");
#endif
#if STEP < 2
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(SW * SH * 4) : "memory");
#endif
  }
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
#endif
  for (int t = threadIdx.x; t < SW * SH; t += blockDim.x) out[t] = smem_buffer[t / SW][t % SW];
}
typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int W = 256, H = 64;
  int *d, *o; cudaMalloc(&d, W * H * 4); cudaMalloc(&o, SW * SH * 4);
  int *h = new int[W * H]; for (int i = 0; i < W * H; ++i) h[i] = i;
  cudaMemcpy(d, h, W * H * 4, cudaMemcpyHostToDevice);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t gdim[2] = {W, H}, gstr[1] = {W * 4};
  cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
#if STEP >= 5
  gdim[0] = 96; gdim[1] = 22; gstr[0] = 384;
  CUresult r = ((Fn)p)(&tm, STEP >= 6 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, (char *)d + 128, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  kernel<<<1, STEP >= 7 ? 256 : 128, SW * SH * 4>>>(tm, STEP >= 8 ? 1 : 0, 0, o);
#else
  CUresult r = ((Fn)p)(&tm, CU_TENSOR_MAP_DATA_TYPE_INT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  kernel<<<1, 128, STEP >= 4 ? SW * SH * 4 : 0>>>(tm, 32, 8, o);
#endif
  cudaError_t e = cudaDeviceSynchronize();
  int g[SW * SH]; cudaMemcpy(g, o, sizeof(g), cudaMemcpyDeviceToHost);
  printf("STEP %d encode %d kernel: %s first %d want %d\n", STEP, (int)r, cudaGetErrorString(e), g[0], 8 * W + 32);
  return 0;
}
