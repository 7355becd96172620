// Plane averages of the output routines, evaluated on the device.
//
// out1d (reference src/output.f90:34-133): profile of a field along idir, averaged over the two other
// directions,  p1d(t) = sum_{a,b} p(a,b,t) w_a(a) w_b(b) * scale,  scale = 1 / (l_a l_b);
// out2d (src/output.f90:135-225): field averaged along idir,  p2d(a,b) = sum_t p(a,b,t) w(t) * scale,
// scale = 1 / l_idir.  The weights are the 1-D spacing arrays the reference passes (dxc/dxf/...,
// ghosted, index lo-1:hi+1).  Each call returns the contribution of one box; the sum over the boxes
// of a block (MPI_ALLREDUCE over comm_block in the reference) and the text file are the host's part
// (snac_b200/output.py).  Only the profile / the plane crosses PCIe instead of the whole field.
#include "common.h"

namespace snac {

namespace {
struct OutArgs {
  const double *p;       // element (0,0,0) of the ghosted field
  long st[3];            // strides of x,y,z
  int n[3];
  int idir, a, b;        // reduced / kept directions, a < b
  const double *wa, *wb; // device copies of the weights (ghosted); out2d: wa = weight along idir
  double scale;
  double *out;
};

// one CTA per index t along idir: sum over the plane (a fastest)
__global__ void __launch_bounds__(256) out1d_kernel(const OutArgs q) {
  const int t = blockIdx.x;
  const long base = (long)(t + 1) * q.st[q.idir];
  const long tot = (long)q.n[q.a] * q.n[q.b];
  double acc = 0.0;
  for (long e = threadIdx.x; e < tot; e += blockDim.x) {
    const int ia = (int)(e % q.n[q.a]) + 1, ib = (int)(e / q.n[q.a]) + 1;
    acc += q.p[base + ia * q.st[q.a] + ib * q.st[q.b]] * __ldg(q.wa + ia) * __ldg(q.wb + ib) * q.scale;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) acc += s[w];
    q.out[t] = acc;
  }
}

// one thread per kept cell (a,b): sum along idir
__global__ void __launch_bounds__(256) out2d_kernel(const OutArgs q) {
  const int ia = blockIdx.x * blockDim.x + threadIdx.x, ib = blockIdx.y;
  if (ia >= q.n[q.a] || ib >= q.n[q.b]) return;
  const long base = (long)(ia + 1) * q.st[q.a] + (long)(ib + 1) * q.st[q.b];
  double acc = 0.0;
  for (int t = 1; t <= q.n[q.idir]; ++t) acc += q.p[base + t * q.st[q.idir]] * __ldg(q.wa + t) * q.scale;
  q.out[ia + (long)ib * q.n[q.a]] = acc;
}
}  // namespace

// which = 1: out1d (host result: n[idir] values), which = 2: out2d (n[a] * n[b] values, a fastest)
void launch_out_reduce(Ctx &ctx, const Grid3 &g, const double *field, int which, int idir, const double *wa, const double *wb,
                       double scale, double *host) {
  if (idir < 0 || idir > 2) throw Error("out1d/out2d: idir must be 0, 1 or 2");
  OutArgs q;
  q.p = field;
  for (int d = 0; d < 3; ++d) { q.st[d] = g.st(d); q.n[d] = g.n[d]; }
  q.idir = idir;
  q.a = idir == 0 ? 1 : 0;
  q.b = idir == 2 ? 1 : 2;
  q.scale = scale;
  const size_t nout = which == 1 ? (size_t)g.n[idir] : (size_t)g.n[q.a] * g.n[q.b];
  const int na = which == 1 ? g.n[q.a] + 2 : g.n[idir] + 2, nb = which == 1 ? g.n[q.b] + 2 : 0;
  DevBuf buf(nout + na + nb);
  SNAC_CUDA(cudaMemcpyAsync(buf.raw + nout, wa, na * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  if (nb) SNAC_CUDA(cudaMemcpyAsync(buf.raw + nout + na, wb, nb * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  q.out = buf.raw;
  q.wa = buf.raw + nout;
  q.wb = buf.raw + nout + na;
  if (which == 1) out1d_kernel<<<g.n[idir], 256, 0, ctx.stream>>>(q);
  else out2d_kernel<<<dim3((g.n[q.a] + 255) / 256, g.n[q.b]), 256, 0, ctx.stream>>>(q);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
  SNAC_CUDA(cudaMemcpyAsync(host, buf.raw, nout * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
}

}  // namespace snac
