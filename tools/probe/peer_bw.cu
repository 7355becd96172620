// Peer-memory bandwidth over NVLink as a function of the segment size of the access pattern (2 GPUs, one process):
// every group of G lanes writes (or reads) a contiguous segment of G*8 bytes; consecutive segments of a CTA are `stride`
// bytes apart at the destination -- the pattern of the wavenumber transposes fused into the y-transforms (fft_radix.cuh:
// 16 pencils x 8 bytes = 128-byte segments, one per wavenumber plane, planes 4 MB apart).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o peer_bw tools/probe/peer_bw.cu && ./peer_bw
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// segment s (of nseg) holds `seg` doubles; segment s of the remote array starts at (s % nplane) * pstride + (s / nplane) * seg
template <bool STORE>
__global__ void k(double *remote, const double *local, double *sink, long nseg, int seg, long nplane, long pstride) {
  const long total = nseg * seg;
  double acc = 0.0;
  for (long t = blockIdx.x * (long)blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
    const long s = t / seg, e = t % seg;
    const long off = (s % nplane) * pstride + (s / nplane) * seg + e;
    if (STORE) remote[off] = local[t]; else acc += remote[off];
  }
  if (!STORE && acc == 1.2345) sink[0] = acc;
  if (STORE) __threadfence_system();
}

int main() {
  int n = 0; CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  const long bytes = 1L << 30, nd = bytes / 8;
  double *loc, *rem, *sink;
  CK(cudaMalloc(&loc, bytes)); CK(cudaMalloc(&sink, 8)); CK(cudaMemset(loc, 0, bytes));
  CK(cudaSetDevice(1)); CK(cudaMalloc(&rem, bytes)); CK(cudaMemset(rem, 0, bytes)); CK(cudaSetDevice(0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int segs[] = {16, 32, 64, 128, 1024};
  const long planes[] = {1, 64, 512};   // 1: fully contiguous; 64 / 512 planes: consecutive segments land 16 MB / 2 MB apart
  for (int store = 1; store >= 0; --store)
    for (long np : planes)
      for (int seg : segs) {
        const long nseg = nd / seg, pstride = nd / np;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          if (store) k<true><<<148 * 8, 256>>>(rem, loc, sink, nseg, seg, np, pstride);
          else k<false><<<148 * 8, 256>>>(rem, loc, sink, nseg, seg, np, pstride);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        printf("%s planes %4ld segment %5d B: %7.1f GB/s\n", store ? "store" : "load ", np, seg * 8, bytes / best / 1e6);
      }
  // the copy engine for comparison
  CK(cudaEventRecord(e0)); CK(cudaMemcpyPeer(rem, 1, loc, 0, bytes)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("cudaMemcpyPeer: %7.1f GB/s\n", bytes / ms / 1e6);
  return 0;
}
