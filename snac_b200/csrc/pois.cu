// K3a: Poisson right-hand side (fillps + updt_rhs), K6: projection + pressure update,
// K7: chkdt/chkdiv reductions.
//
// fillps   : reference src/fillps.f90:7-54       updt_rhs      : src/bound.f90:543-585
// correc   : src/correc.f90:7-48                 updt_pressure : src/updt_pressure.f90:7-47
// chkdt    : src/chkdt.f90:8-56                  chkdiv        : src/chkdiv.f90:8-44
// init_bc_rhs (src/solver.f90:44-157) is folded into updt_rhs: its planes are separable,
// rhs(a,b) = coef * w_a(a) * w_b(b), so only one coefficient per face is stored.
#include "common.h"

namespace snac {

struct PoisArgs {
  BoxGeom g;
  const double *u, *v, *w;
  double *rhs;
  const double *wq[3];  // ghosted weights: cell width, or 1 along the transform axis
  const double *gq[3];  // 1, or 1/width along the transform axis
  double cb[3][2];      // boundary RHS coefficient per face (0: none)
  double rdt;
  int do_fill, do_bc;   // do_fill = 2: rhs = alphai * weight * fld (Helmholtz right-hand side)
  const double *fld;
  double alphai;
  int nhi[3];           // range on which the boundary terms are added (shortened for staggered unknowns)
};

__global__ void __launch_bounds__(256) fillps_kernel(const PoisArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > a.g.n[0] || j > a.g.n[1] || k > a.g.n[2]) return;
  const long sy = a.g.sy, sz = a.g.sz, c = i + sy * j + sz * k;
  const double wx = __ldg(a.wq[0] + i), wy = __ldg(a.wq[1] + j), wz = __ldg(a.wq[2] + k);
  double val;
  if (a.do_fill == 2) {
    val = a.alphai * wx * wy * wz * a.fld[c];   // src/solve_helmholtz.f90:121-129
  } else if (a.do_fill) {
    const double du = a.u[c] - a.u[c - 1], dv = a.v[c] - a.v[c - sy], dw = a.w[c] - a.w[c - sz];
    val = (dw * __ldg(a.gq[2] + k) * wx * wy + dv * __ldg(a.gq[1] + j) * wx * wz +
           du * __ldg(a.gq[0] + i) * wy * wz) * a.rdt;
  } else {
    val = a.rhs[c];
  }
  if (a.do_bc) {
    // updt_rhs acts on lo:hi of the (possibly shortened) unknown range (src/bound.f90:543-585)
    const bool in = i <= a.nhi[0] && j <= a.nhi[1] && k <= a.nhi[2];
    if (in) {
      if (i == 1) val += a.cb[0][0] * wy * wz;
      if (i == a.nhi[0]) val += a.cb[0][1] * wy * wz;
      if (j == 1) val += a.cb[1][0] * wx * wz;
      if (j == a.nhi[1]) val += a.cb[1][1] * wx * wz;
      if (k == 1) val += a.cb[2][0] * wx * wy;
      if (k == a.nhi[2]) val += a.cb[2][1] * wx * wy;
    }
  }
  a.rhs[c] = val;
}

// rows of 64 threads per block: measured 4 / 2 / 1 on TGV-256: fillps 0.42 / 0.40 / 0.42 ms, correc 0.84 / 0.81 / 0.83 ms per step
#ifndef SNAC_POIS_BY
#define SNAC_POIS_BY 2
#endif
static dim3 pick_block(const Grid3 &g) {
  if (g.n[0] <= 8) return dim3(8, 8, 4);
  if (g.n[0] <= 32) return dim3(32, 8, 1);
  return dim3(64, SNAC_POIS_BY, 1);
}
static dim3 grid_for(const dim3 &blk, int n0, int n1, int n2) {
  return dim3((n0 + blk.x - 1) / blk.x, (n1 + blk.y - 1) / blk.y, (n2 + blk.z - 1) / blk.z);
}

static void fill_common(Ctx &ctx, Box &b, double dtrk, int do_fill, int do_bc) {
  if (do_fill == 1 && launch_fillps_tma(ctx, b, dtrk, do_bc != 0)) return;
  PoisArgs a;
  a.g = b.geom();
  a.u = b.u[ctx.par].p; a.v = b.v[ctx.par].p; a.w = b.w[ctx.par].p;
  a.rhs = b.rhs.p;
  for (int d = 0; d < 3; ++d) {
    a.wq[d] = b.wq[d].p();
    a.gq[d] = b.gq[d].p();
    a.cb[d][0] = b.cbp[d][0];
    a.cb[d][1] = b.cbp[d][1];
  }
  a.rdt = do_fill ? 1.0 / dtrk : 0.0;
  a.do_fill = do_fill;
  a.do_bc = do_bc;
  a.fld = nullptr; a.alphai = 0.0;
  for (int d = 0; d < 3; ++d) a.nhi[d] = b.g.n[d];
  dim3 blk = pick_block(b.g), grd = grid_for(blk, b.g.n[0], b.g.n[1], b.g.n[2]);
  fillps_kernel<<<grd, blk, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

// right-hand side of an implicit-diffusion Helmholtz problem into the box's rhs array
void launch_helm_rhs(Ctx &ctx, Box &b, const double *fld, const double *const wq[3], const double cb[3][2], const int nhi[3],
                     double alphai) {
  PoisArgs a;
  a.g = b.geom();
  a.u = a.v = a.w = nullptr;
  a.rhs = b.rhs.p;
  for (int d = 0; d < 3; ++d) {
    a.wq[d] = wq[d]; a.gq[d] = nullptr;
    a.cb[d][0] = cb[d][0]; a.cb[d][1] = cb[d][1];
    a.nhi[d] = nhi[d];
  }
  a.rdt = 0.0; a.do_fill = 2; a.do_bc = 1; a.fld = fld; a.alphai = alphai;
  dim3 blk = pick_block(b.g), grd = grid_for(blk, b.g.n[0], b.g.n[1], b.g.n[2]);
  fillps_kernel<<<grd, blk, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

void launch_fillps(Ctx &ctx, Box &b, double dtrk) { fill_common(ctx, b, dtrk, 1, 0); }
void launch_updt_rhs(Ctx &ctx, Box &b) { fill_common(ctx, b, 1.0, 0, 1); }
void launch_fillps_bc(Ctx &ctx, Box &b, double dtrk) { fill_common(ctx, b, dtrk, 1, 1); }

// Boundary RHS coefficients of a centred (pressure/scalar) problem: init_bc_rhs restricted to
// is_centered = (T,T,T) (src/solver.f90:71-94,125).
void compute_bc_rhs(Ctx &ctx, Box &b) {
  const int fa = ctx.prm.fft_axis - 1;
  for (int d = 0; d < 3; ++d)
    for (int ib = 0; ib < 2; ++ib) {
      b.cbp[d][ib] = 0.0;
      if (!b.info.is_bound[d][ib] || ctx.prm.periods[d] != 0) continue;
      if (ib == 1 && b.g.n[d] == 1) continue;  // the reference loop visits i==lo only (ib=0)
      const char c = b.info.cbcpre[d][ib];
      const double bc = b.info.bcpre[d][ib];
      double factor = 0.0;
      if (c == 'N') factor = (ib == 0 ? 1.0 : -1.0) * b.info.dl[d][ib] * bc;
      else if (c == 'D') factor = -2.0 * bc;
      const int i = (ib == 0) ? 1 : b.g.n[d];
      const double d1 = b.h_dc[d][i - (1 - ib)], d2 = b.h_df[d][i];
      const double wn = (d == fa) ? 1.0 : d2;
      b.cbp[d][ib] = factor * wn / (d1 * d2);
    }
}

// ------------------------------------------------------------------------------------------
struct CorrArgs {
  BoxGeom g;
  double *u, *v, *w, *p;
  const double *pp;
  double dtrk, alpha;
  int do_correc, do_updt, impdiff;
};

__global__ void __launch_bounds__(256) correc_kernel(const CorrArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // full ghosted extents 0..n+1
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int k = blockIdx.z * blockDim.z + threadIdx.z;
  const int n0 = a.g.n[0], n1 = a.g.n[1], n2 = a.g.n[2];
  if (i > n0 + 1 || j > n1 + 1 || k > n2 + 1) return;
  const long sy = a.g.sy, sz = a.g.sz, c = i + sy * j + sz * k;
  const double p0 = a.pp[c];
  const bool interior = (i >= 1 && i <= n0 && j >= 1 && j <= n1 && k >= 1 && k <= n2);
  double pip = 0, pjp = 0, pkp = 0;
  if (i <= n0) pip = a.pp[c + 1];
  if (j <= n1) pjp = a.pp[c + sy];
  if (k <= n2) pkp = a.pp[c + sz];
  if (a.do_correc) {  // ranges include the ghost layers (src/correc.f90:23-47)
    if (i <= n0) a.u[c] -= a.dtrk * (pip - p0) * __ldg(a.g.rdc[0] + i);
    if (j <= n1) a.v[c] -= a.dtrk * (pjp - p0) * __ldg(a.g.rdc[1] + j);
    if (k <= n2) a.w[c] -= a.dtrk * (pkp - p0) * __ldg(a.g.rdc[2] + k);
  }
  if (a.do_updt && interior) {
    double val = a.p[c] + p0;
    if (a.impdiff) {  // src/updt_pressure.f90:30-46
      const double pim = a.pp[c - 1], pjm = a.pp[c - sy], pkm = a.pp[c - sz];
      val += a.alpha *
             (((pip - p0) * __ldg(a.g.rdc[0] + i) - (p0 - pim) * __ldg(a.g.rdc[0] + i - 1)) * __ldg(a.g.rdf[0] + i) +
              ((pjp - p0) * __ldg(a.g.rdc[1] + j) - (p0 - pjm) * __ldg(a.g.rdc[1] + j - 1)) * __ldg(a.g.rdf[1] + j) +
              ((pkp - p0) * __ldg(a.g.rdc[2] + k) - (p0 - pkm) * __ldg(a.g.rdc[2] + k - 1)) * __ldg(a.g.rdf[2] + k));
    }
    a.p[c] = val;
  }
}

void launch_correc_updt(Ctx &ctx, Box &b, double dtrk, double alpha, bool do_correc, bool do_updt) {
  if (launch_correc_tma(ctx, b, dtrk, alpha, do_correc, do_updt)) return;
  CorrArgs a;
  a.g = b.geom();
  a.u = b.u[ctx.par].p; a.v = b.v[ctx.par].p; a.w = b.w[ctx.par].p;
  a.p = b.p.p;
  a.pp = b.pp_ptr();
  a.dtrk = dtrk; a.alpha = alpha;
  a.do_correc = do_correc; a.do_updt = do_updt; a.impdiff = ctx.prm.impdiff;
  dim3 blk = pick_block(b.g), grd = grid_for(blk, b.g.n[0] + 2, b.g.n[1] + 2, b.g.n[2] + 2);
  correc_kernel<<<grd, blk, 0, ctx.stream>>>(a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
struct ChkArgs {
  BoxGeom g;
  const double *u, *v, *w;
  double rvol;
  double *out;  // [0] dti max, [1] divtot, [2] divmax
};

__device__ inline void atomic_max_pos(double *addr, double v) {  // v >= 0
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

// A block owns a tile of columns and a chunk of kz planes and accumulates in registers: three atomics per block and
// chunk (one block per 128 cells, i.e. 1.5 M atomics on three addresses for a 1024 x 128 x 512 box, ran at 0.7 TB/s).
__global__ void __launch_bounds__(256) chk_kernel(const ChkArgs a, int kz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k0 = blockIdx.z * kz + 1, k1 = min(k0 + kz - 1, a.g.n[2]);
  double dti = 0.0, divs = 0.0, divm = 0.0;
  if (i <= a.g.n[0] && j <= a.g.n[1]) {
    const long sy = a.g.sy, sz = a.g.sz;
    const double rdxc = __ldg(a.g.rdc[0] + i), rdyc = __ldg(a.g.rdc[1] + j);
    const double rdxf = __ldg(a.g.rdf[0] + i), rdyf = __ldg(a.g.rdf[1] + j);
    const double dfxy = __ldg(a.g.df[0] + i) * __ldg(a.g.df[1] + j);
    for (int k = k0; k <= k1; ++k) {
      const long c = i + sy * j + sz * k;
      const double *u = a.u + c, *v = a.v + c, *w = a.w + c;
      const double rdzc = __ldg(a.g.rdc[2] + k), rdzf = __ldg(a.g.rdf[2] + k);
      const double u0 = u[0], v0 = v[0], w0 = w[0];
      const double uim = u[-1], vjm = v[-sy], wkm = w[-sz];
      const double ux = fabs(u0);
      const double vx = 0.25 * fabs(v0 + vjm + v[1] + v[1 - sy]);
      const double wx = 0.25 * fabs(w0 + wkm + w[1] + w[1 - sz]);
      const double dtix = ux * rdxc + vx * rdyf + wx * rdzf;
      const double uy = 0.25 * fabs(u0 + u[sy] + u[sy - 1] + uim);
      const double vy = fabs(v0);
      const double wy = 0.25 * fabs(w0 + w[sy] + w[sy - sz] + wkm);
      const double dtiy = uy * rdxf + vy * rdyc + wy * rdzf;
      const double uz = 0.25 * fabs(u0 + uim + u[sz - 1] + u[sz]);
      const double vz = 0.25 * fabs(v0 + vjm + v[sz - sy] + v[sz]);
      const double wz = fabs(w0);
      const double dtiz = uz * rdxf + vz * rdyf + wz * rdzc;
      dti = fmax(dti, fmax(dtix, fmax(dtiy, dtiz)));
      const double div = (w0 - wkm) * rdzf + (v0 - vjm) * rdyf + (u0 - uim) * rdxf;
      divm = fmax(divm, fabs(div));
      divs += div * dfxy * __ldg(a.g.df[2] + k) * a.rvol;
    }
  }
  // block reduction
  for (int o = 16; o > 0; o >>= 1) {
    dti = fmax(dti, __shfl_xor_sync(0xffffffffu, dti, o));
    divm = fmax(divm, __shfl_xor_sync(0xffffffffu, divm, o));
    divs += __shfl_xor_sync(0xffffffffu, divs, o);
  }
  __shared__ double s[3][8];
  const int tid = threadIdx.x + blockDim.x * threadIdx.y;
  const int nw = (blockDim.x * blockDim.y + 31) / 32;
  if ((tid & 31) == 0) { s[0][tid >> 5] = dti; s[1][tid >> 5] = divs; s[2][tid >> 5] = divm; }
  __syncthreads();
  if (tid == 0) {
    for (int q = 1; q < nw; ++q) { dti = fmax(dti, s[0][q]); divs += s[1][q]; divm = fmax(divm, s[2][q]); }
    atomic_max_pos(a.out + 0, dti);
    atomicAdd(a.out + 1, divs);
    atomic_max_pos(a.out + 2, divm);
  }
}

void launch_chk(Ctx &ctx, Box &b, double *d_out) {
  ChkArgs a;
  a.g = b.geom();
  a.u = b.u[ctx.par].p; a.v = b.v[ctx.par].p; a.w = b.w[ctx.par].p;
  a.rvol = 1.0 / ctx.prm.vol_all;
  a.out = d_out;
  dim3 blk = pick_block(b.g);
  blk.y *= blk.z; blk.z = 1;
  const int kz = 16;
  dim3 grd = grid_for(blk, b.g.n[0], b.g.n[1], 1);
  grd.z = (b.g.n[2] + kz - 1) / kz;
  chk_kernel<<<grd, blk, 0, ctx.stream>>>(a, kz);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

}  // namespace snac
