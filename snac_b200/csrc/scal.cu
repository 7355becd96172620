// K1s: scalar transport (reference rk_scal, src/rk.f90:130-204; scal_d / scal_a, src/scal.f90:89-154;
// bulk_mean_s / bulk_forcing_s, src/scal.f90:156-199).
// Three passes: (1) new right-hand side from the old field (stencil, reads neighbours: written to a
// scratch array), (2) pointwise RK update + volume sum for the bulk forcing, (3) pointwise
// completion (implicit-diffusion half step, forcing).
#include "common.h"

namespace snac {

struct ScalArgs {
  BoxGeom g;
  const double *u, *v, *w;
  double *s, *ds, *rk, *dd;
  double alpha, f1, f2, f12, src, f;
  int impdiff, forced;
  double *sum;
};

static dim3 sblock(const Grid3 &g) {
  if (g.n[0] <= 8) return dim3(8, 8, 4);
  if (g.n[0] <= 32) return dim3(32, 8, 1);
  return dim3(64, 4, 1);
}
static dim3 sgrid(const dim3 &blk, const Grid3 &g) {
  return dim3((g.n[0] + blk.x - 1) / blk.x, (g.n[1] + blk.y - 1) / blk.y, (g.n[2] + blk.z - 1) / blk.z);
}

__global__ void __launch_bounds__(256) scal_rhs_kernel(const ScalArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > a.g.n[0] || j > a.g.n[1] || k > a.g.n[2]) return;
  const long sy = a.g.sy, sz = a.g.sz, c = i + sy * j + sz * k;
  const double *s = a.s + c;
  const double s0 = s[0], sip = s[1], sim = s[-1], sjp = s[sy], sjm = s[-sy], skp = s[sz], skm = s[-sz];
  const double rdxf = __ldg(a.g.rdf[0] + i), rdyf = __ldg(a.g.rdf[1] + j), rdzf = __ldg(a.g.rdf[2] + k);
  // diffusion (src/scal.f90:104-117)
  const double dd = a.alpha * (((sip - s0) * __ldg(a.g.rdc[0] + i) - (s0 - sim) * __ldg(a.g.rdc[0] + i - 1)) * rdxf +
                               ((sjp - s0) * __ldg(a.g.rdc[1] + j) - (s0 - sjm) * __ldg(a.g.rdc[1] + j - 1)) * rdyf +
                               ((skp - s0) * __ldg(a.g.rdc[2] + k) - (s0 - skm) * __ldg(a.g.rdc[2] + k - 1)) * rdzf);
  // advection (src/scal.f90:138-150)
  const double aa = (-0.5 * (sip + s0) * a.u[c] + 0.5 * (sim + s0) * a.u[c - 1]) * rdxf +
                    (-0.5 * (sjp + s0) * a.v[c] + 0.5 * (sjm + s0) * a.v[c - sy]) * rdyf +
                    (-0.5 * (skp + s0) * a.w[c] + 0.5 * (skm + s0) * a.w[c - sz]) * rdzf;
  a.rk[c] = a.impdiff ? aa : dd + aa;
  if (a.impdiff) a.dd[c] = dd;
}

__global__ void __launch_bounds__(256) scal_update_kernel(const ScalArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  double part = 0.0;
  if (i <= a.g.n[0] && j <= a.g.n[1] && k <= a.g.n[2]) {
    const long c = i + a.g.sy * j + a.g.sz * k;
    const double rk = a.rk[c];
    double sn = a.s[c] + a.f1 * rk + a.f2 * a.ds[c] + a.f12 * a.src;
    if (a.impdiff) sn += a.f12 * a.dd[c];
    a.ds[c] = rk;
    a.s[c] = sn;
    if (a.forced) part = sn * __ldg(a.g.df[0] + i) * __ldg(a.g.df[1] + j) * __ldg(a.g.df[2] + k);
  }
  if (a.forced) {
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ double sh[8];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    if ((tid & 31) == 0) sh[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) / 32;
      for (int q = 1; q < nw; ++q) part += sh[q];
      atomicAdd(a.sum, part);
    }
  }
}

__global__ void __launch_bounds__(256) scal_finish_kernel(const ScalArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > a.g.n[0] || j > a.g.n[1] || k > a.g.n[2]) return;
  const long c = i + a.g.sy * j + a.g.sz * k;
  double sn = a.s[c];
  if (a.impdiff) sn -= 0.5 * a.f12 * a.dd[c];
  if (a.forced) sn += a.f;
  a.s[c] = sn;
}

static ScalArgs sargs(Ctx &ctx, Box &b, int is) {
  ScalArgs a;
  a.g = b.geom();
  a.u = b.u[ctx.par].p; a.v = b.v[ctx.par].p; a.w = b.w[ctx.par].p;
  a.s = b.s[is].p; a.ds = b.ds[is].p; a.rk = b.scr0.p; a.dd = b.scr1.p;
  a.alpha = 1.0 / ctx.prm.alphai[is];
  a.f1 = a.f2 = a.f12 = 0.0;
  a.src = ctx.prm.ssource[is];
  a.f = 0.0;
  a.impdiff = ctx.prm.impdiff;
  a.forced = ctx.prm.is_sforced[is];
  a.sum = nullptr;
  return a;
}

void launch_scal_rhs(Ctx &ctx, Box &b, int is, double alpha) {
  ScalArgs a = sargs(ctx, b, is);
  a.alpha = alpha;
  dim3 blk = sblock(b.g);
  scal_rhs_kernel<<<sgrid(blk, b.g), blk, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}
void launch_scal_update(Ctx &ctx, Box &b, int is, double f1, double f2, double f12, double *d_sum) {
  ScalArgs a = sargs(ctx, b, is);
  a.f1 = f1; a.f2 = f2; a.f12 = f12; a.sum = d_sum;
  dim3 blk = sblock(b.g);
  scal_update_kernel<<<sgrid(blk, b.g), blk, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}
void launch_scal_finish(Ctx &ctx, Box &b, int is, double f12, double f) {
  ScalArgs a = sargs(ctx, b, is);
  a.f12 = f12; a.f = f;
  dim3 blk = sblock(b.g);
  scal_finish_kernel<<<sgrid(blk, b.g), blk, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

}  // namespace snac
