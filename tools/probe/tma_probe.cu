// stand-alone probe of the TMA path used by mom_tma.cu: one 3-D FP64 box load per slot, mbarrier completion
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
constexpr int BX = 34, BY = 10, BUF = ((BX * BY * 8 + 127) / 128) * 128 / 8;
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, double *out, int c0, int c1, int c2, int nbytes, int variant) {
  extern __shared__ __align__(128) double ring[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (variant & 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!(variant & 2)) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(smem_u32(ring)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(smem_u32(ring)), "l"(&tm), "r"(smem_u32(&bar)), "r"(c0), "r"(c1) : "memory");
    if (variant & 2) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nbytes) : "memory");
  }
  asm volatile(
      "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int t = threadIdx.x; t < nbytes / 8; t += blockDim.x) out[t] = ring[t];
}
typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  const int promo = 0, variant = argc > 1 ? atoi(argv[1]) : 0;
  const int n1 = 40, n2 = 20, n3 = 6;
  const long sy = 48, sz = sy * (n2 + 2);
  const size_t elems = 16 + sz * (n3 + 2);
  std::vector<double> h(elems);
  for (size_t i = 0; i < elems; ++i) h[i] = (double)i;
  double *d, *o;
  cudaMalloc(&d, elems * 8); cudaMalloc(&o, BX * BY * 8 * 2);
  cudaMemcpy(d, h.data(), elems * 8, cudaMemcpyHostToDevice);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %d %d %p\n", (int)e, (int)q, p);
  struct Cfg { int rank; CUtensorMapDataType dt; int es; int bx, by; int off; const char *name; };
  const Cfg cfgs[] = {{2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 32, 8, 16, "f32 2d 32x8 off16"},
                      {2, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 32, 8, 16, "f64 2d 32x8 off16"},
                      {3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 32, 8, 16, "f64 3d 32x8 off16"},
                      {3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 34, 10, 16, "f64 3d 34x10 off16"},
                      {3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 34, 10, 14, "f64 3d 34x10 off14"},
                      {3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 8, 36, 10, 14, "f64 3d 36x10 off14"},
                      {3, CU_TENSOR_MAP_DATA_TYPE_INT64, 8, 34, 10, 14, "i64 3d 34x10 off14"},
                      {3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 34 * 8, 10, 14, "u8 3d 272x10 off14 (bytes)"}};
  for (const Cfg &c : cfgs) {
    CUtensorMap tm;
    const cuuint64_t es = c.es;
    const cuuint64_t gdim[3] = {(cuuint64_t)sy * 8 / es, (cuuint64_t)(n2 + 2), (cuuint64_t)(n3 + 2)};
    const cuuint64_t gstr[2] = {(cuuint64_t)sy * 8, (cuuint64_t)sz * 8};
    const cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, 1}, estr[3] = {1, 1, 1};
    CUresult r = ((Fn)p)(&tm, c.dt, c.rank, d + c.off, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int nbytes = c.bx * c.by * c.es;
    cudaMemset(o, 0, BX * BY * 8 * 2);
    if (c.rank == 3) {
      cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * BUF * 8);
      probe<3><<<1, 256, 8 * BUF * 8>>>(tm, o, c.es == 1 ? 8 : 1, 0, 2, nbytes, variant);
    } else {
      cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * BUF * 8);
      probe<2><<<1, 256, 8 * BUF * 8>>>(tm, o, 1, 0, 0, nbytes, variant);
    }
    e = cudaDeviceSynchronize();
    printf("%-32s encode %d kernel: %s\n", c.name, (int)r, cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("context is dead after this\n"); break; }
  }
  return 0;
}
