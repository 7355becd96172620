// K4: batched geometric multigrid for the pressure Poisson / Helmholtz problems.
//
// Replaces the HYPRE struct solves of the reference: solve_n_helmholtz_2d (src/solver.f90:843-922,
// one PFMG solve per wavenumber), solve_helmholtz (src/solver.f90:524-573, one 3-D solve), and the
// matrix set-up init_matrix_2d/init_matrix_3d/init_n_2d_matrices (src/solver.f90:158-372,687-842,
// 981-1020) with create_solver/setup_solver (src/solver.f90:373-484).  HYPRE is an external,
// unpinned dependency of the reference; what is reproduced here is the linear system it is given
// (coefficients rebuilt from the 1-D spacing arrays, never stored per cell) and its stopping rule
// (relative residual < hypre_tol, checked on the residual after pre-relaxation as in PFMG; a zero
// right-hand side returns a zero solution).
//
// The operator is a tensor product: for every direction d that is coupled,
//   (A x)(i,j,k) = sum_d W_notd * ( am_d[i] x(i-1) + ap_d[i] x(i+1) + ad_d[i] x(i) )
//                  + (alpha + lambda_q) * W * x(i,j,k),
// where W is the product of the cell widths of the coupled directions.  With FFT_AXIS on, the
// transform direction is uncoupled and indexes the batch of independent 2-D problems (one per
// wavenumber q, eigenvalue lambda_q); they are all advanced by the same kernels.  Boundary
// conditions are folded into ad/am/ap exactly as init_matrix_* does (src/solver.f90:722-743,
// 788-823); block interfaces and periodic wraps go through the ghost layer of x, refreshed by the
// same halo machinery as the flow fields.
//
// Cycle: V(npre,npost) with red-black Gauss-Seidel, cell-centred coarsening by pairs with
// re-discretised coarse operators (coarse cell width = sum of its two children), residual
// summation for restriction and piecewise-constant prolongation.  Directions are coarsened only
// while their mesh width is within a factor 1.5 of the finest one (semi-coarsening for anisotropic
// grids).  A box of odd size keeps its last cell unpaired, so coarsening continues down to one
// cell per box and direction.
#include <algorithm>
#include <cmath>
#include <math_constants.h>

#include "common.h"
#include "mg_dev.cuh"

namespace snac {

// ------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mg_smooth_kernel(const MGKernelBox B, int color, double alpha) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > B.n[0] || j > B.n[1] || k > B.n[2]) return;
  const int par = (B.coupled[0] ? i : 0) + (B.coupled[1] ? j : 0) + (B.coupled[2] ? k : 0);
  if ((par & 1) != color) return;
  const long c = i + B.sy * j + B.sz * k;
  const double wx = __ldg(B.w[0] + i), wy = __ldg(B.w[1] + j), wz = __ldg(B.w[2] + k);
  const double diag = mg_diag(B, i, j, k, wx, wy, wz, alpha);
  if (diag == 0.0) return;
  const double off = mg_offdiag(B, B.x, c, i, j, k, wx, wy, wz);
  B.x[c] = (B.b[c] - off) / diag;
}

// fine residual summed over the children of each coarse cell -> coarse right-hand side; optionally
// accumulates per-plane sums of r^2 and b^2 (norms[q], norms[nq+q]).
__global__ void __launch_bounds__(256) mg_restrict_kernel(const MGKernelBox F, double *cb, long csy, long csz,
                                                          int cn0, int cn1, int cn2, int c0, int c1, int c2,
                                                          double alpha, double *norms, int nq, int want_b) {
  const int I = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int J = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int K = blockIdx.z * blockDim.z + threadIdx.z + 1;
  const bool active = (I <= cn0 && J <= cn1 && K <= cn2);
  double rsum = 0.0, r2 = 0.0, b2 = 0.0;
  if (active) {
    // children of a coarse cell: the pair (2I-1, 2I); an odd fine size leaves the last one single
    const int i0 = c0 ? 2 * I - 1 : I, i1 = c0 ? min(2 * I, F.n[0]) : I;
    const int j0 = c1 ? 2 * J - 1 : J, j1 = c1 ? min(2 * J, F.n[1]) : J;
    const int k0 = c2 ? 2 * K - 1 : K, k1 = c2 ? min(2 * K, F.n[2]) : K;
    for (int k = k0; k <= k1; ++k)
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          const long c = i + F.sy * j + F.sz * k;
          const double wx = __ldg(F.w[0] + i), wy = __ldg(F.w[1] + j), wz = __ldg(F.w[2] + k);
          const double bb = F.b[c];
          const double r = bb - (mg_diag(F, i, j, k, wx, wy, wz, alpha) * F.x[c] + mg_offdiag(F, F.x, c, i, j, k, wx, wy, wz));
          rsum += r;
          r2 += r * r;
          b2 += bb * bb;
        }
    if (cb) cb[I + csy * J + csz * K] = rsum;
  }
  if (norms) {
    // threads of a block share the plane index unless the uncoupled direction is x
    if (F.fft_dir == 0) {
      if (active) {
        atomicAdd(norms + (I - 1), r2);
        if (want_b) atomicAdd(norms + nq + (I - 1), b2);
      }
      return;
    }
    for (int o = 16; o > 0; o >>= 1) {
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    __shared__ double s[2][8];
    const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
    const int nw = (blockDim.x * blockDim.y * blockDim.z + 31) / 32;
    if ((tid & 31) == 0) { s[0][tid >> 5] = r2; s[1][tid >> 5] = b2; }
    __syncthreads();
    if (tid == 0) {
      for (int q = 1; q < nw; ++q) { r2 += s[0][q]; b2 += s[1][q]; }
      const int q = F.fft_dir < 0 ? 0 : (F.fft_dir == 1 ? (int)(blockIdx.y * blockDim.y) : (int)(blockIdx.z * blockDim.z));
      atomicAdd(norms + q, r2);
      if (want_b) atomicAdd(norms + nq + q, b2);
    }
  }
}

__global__ void __launch_bounds__(256) mg_prolong_kernel(const MGKernelBox F, const double *cx, long csy, long csz,
                                                         int c0, int c1, int c2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > F.n[0] || j > F.n[1] || k > F.n[2]) return;
  const int I = c0 ? (i + 1) >> 1 : i, J = c1 ? (j + 1) >> 1 : j, K = c2 ? (k + 1) >> 1 : k;
  F.x[i + F.sy * j + F.sz * k] += cx[I + csy * J + csz * K];
}

// zero the planes whose right-hand side vanishes (HYPRE returns x = 0 for b = 0)
__global__ void __launch_bounds__(256) mg_zero_planes_kernel(const MGKernelBox B, const double *norms, int nq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > B.n[0] || j > B.n[1] || k > B.n[2]) return;
  const int q = B.fft_dir < 0 ? 0 : (B.fft_dir == 0 ? i - 1 : (B.fft_dir == 1 ? j - 1 : k - 1));
  if (norms[nq + q] == 0.0) B.x[i + B.sy * j + B.sz * k] = 0.0;
}

__global__ void mg_fill_along_kernel(double *x, long sy, long sz, int n0, int n1, int n2, int dir, const double *w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > n0 || j > n1 || k > n2) return;
  x[i + sy * j + sz * k] = w[dir == 0 ? i : (dir == 1 ? j : k)];
}

// ------------------------------------------------------------------------------------------
// plane-local fast path (batched 2-D problems whose planes are not split over boxes): the fused
// fine-level kernels live in mg_ring.cu, the one-CTA-per-plane kernel for every level that fits in
// shared memory in mg_coarse.cu; what follows are the per-plane stopping flags they share.
// ------------------------------------------------------------------------------------------
// ---- per-wavenumber stopping test on the device ------------------------------------------------
// norms = [sum r^2 | sum b^2 | relative residual] per plane.  A plane leaves the iteration when its
// relative residual (after pre-relaxation, as in PFMG) is below the tolerance (from the second
// cycle on), when its right-hand side is zero (x = 0), or at maxiter; the kernels of later cycles
// skip it.  where[q]: 0 result in the main x buffer, 1 in the second buffer, 2 identically zero.
__global__ void mg_flags_init_kernel(int *active, int *where, int *iters, int *count, int nq) {
  for (int q = threadIdx.x; q < nq; q += blockDim.x) { active[q] = 1; where[q] = 0; iters[q] = 0; }
  if (threadIdx.x == 0) *count = nq;
}

__global__ void mg_flags_kernel(double *norms, int nq, double tol2, int it, int maxiter, int in_alt, int *active, int *where,
                                int *iters, int *count) {
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    if (!active[q]) continue;
    const double r2 = norms[q], b2 = norms[nq + q];
    int act = 1;
    if (b2 == 0.0) {
      act = 0; where[q] = 2; iters[q] = 0; norms[2 * nq + q] = 0.0;
    } else {
      const double rel2 = r2 / b2;
      norms[2 * nq + q] = sqrt(rel2);
      iters[q] = it;
      if ((it > 0 && rel2 < tol2) || it >= maxiter) { act = 0; where[q] = in_alt ? 1 : 0; }
    }
    active[q] = act;
    if (act) { atomicAdd(&cnt, 1); norms[q] = 0.0; }
  }
  __syncthreads();
  if (threadIdx.x == 0) *count = cnt;
}

// zero_only: the buffers have traded places, planes with where == 1 are already in x0
__global__ void mg_finalize_kernel(double *x0, const double *x1, long sy, long sz, int n1, int n2, const int *where, int zero_only) {
  const int q = blockIdx.z, w = where[q];
  if (w == 0 || (zero_only && w == 1)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1, j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  if (i > n1 || j > n2) return;
  const long o = (long)(q + 1) * sz + i + sy * j;
  x0[o] = (w == 1) ? x1[o] : 0.0;
}

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
// smoothing sweeps per leg of the V-cycle: snac_params, else SNAC_MG_NSWEEP, else the solver's default
static void mg_sweeps(const Ctx &ctx, const MGSolver &S, int &npre, int &npost) {
  int def = S.default_sweeps;
  npre = ctx.prm.mg_npre > 0 ? ctx.prm.mg_npre : def;
  npost = ctx.prm.mg_npost > 0 ? ctx.prm.mg_npost : def;
}

static dim3 mg_block(const Grid3 &g, int fft_dir) {
  // threads of a block must share the plane index along z when that is the batch direction
  dim3 b;
  if (g.n[0] <= 8) b = dim3(8, 8, 4);
  else if (g.n[0] <= 32) b = dim3(32, 8, 1);
  else b = dim3(64, 4, 1);
  if (fft_dir == 1) { b.z *= b.y; b.y = 1; if (b.z > 64) { b.z = 64; } }
  if (fft_dir == 2) { b.y *= b.z; b.z = 1; }
  return b;
}
static dim3 mg_grid(const dim3 &blk, int n0, int n1, int n2) {
  return dim3((n0 + blk.x - 1) / blk.x, (n1 + blk.y - 1) / blk.y, (n2 + blk.z - 1) / blk.z);
}

static MGKernelBox kbox(const MGSolver &S, const MGLevelBox &L, int lb, const bool coupled[3]) {
  MGKernelBox B;
  for (int d = 0; d < 3; ++d) {
    B.n[d] = L.g.n[d];
    B.am[d] = L.am[d].p(); B.ap[d] = L.ap[d].p(); B.ad[d] = L.ad[d].p(); B.w[d] = L.w[d].p();
    B.coupled[d] = coupled[d];
  }
  B.sy = L.g.sy; B.sz = L.g.sz;
  B.x = L.xp; B.b = L.bp;
  B.fft_dir = S.fft_dir;
  B.lam = S.fft_dir >= 0 ? S.lam[lb].p() : nullptr;
  return B;
}

// builds am/ap/ad/w of one direction of one box from ghosted spacings d1 (between unknowns) and
// d2 (cell widths), folding physical boundary conditions into the diagonal
static void build_dir(MGLevelBox &L, int d, bool coupled, const std::vector<double> &d1, const std::vector<double> &d2,
                      const double sgn[2], const bool phys[2], int qqq = 0) {
  const int n = L.g.n[d];
  std::vector<double> am(n + 2, 0.0), ap(n + 2, 0.0), ad(n + 2, 0.0), w(n + 2, 1.0);
  if (coupled) {
    for (int i = 0; i <= n + 1; ++i) w[i] = d2[i];
    for (int i = 1; i <= n; ++i) {
      am[i] = 1.0 / d1[i - 1 + qqq];   // src/solver.f90:289-297 (index shift of staggered unknowns)
      ap[i] = 1.0 / d1[i + qqq];
      ad[i] = -(am[i] + ap[i]);
    }
    if (phys[0]) { ad[1] += sgn[0] * am[1]; am[1] = 0.0; }
    if (phys[1]) { ad[n] += sgn[1] * ap[n]; ap[n] = 0.0; }
  }
  L.h_w[d] = w;
  L.am[d].upload(am); L.ap[d].upload(ap); L.ad[d].upload(ad); L.w[d].upload(w);
  std::vector<double> sam(n + 2, 0.0), sap(n + 2, 0.0), se(n + 2, 0.0), srw(n + 2, 1.0);
  for (int i = 1; i <= n; ++i) {
    srw[i] = 1.0 / w[i];
    sam[i] = am[i] * srw[i]; sap[i] = ap[i] * srw[i]; se[i] = ad[i] * srw[i];
  }
  L.sam[d].upload(sam); L.sap[d].upload(sap); L.se[d].upload(se); L.srw[d].upload(srw);
  std::vector<double> tab(8 * (size_t)(n + 2), 0.0);
  for (int i = 0; i <= n + 1; ++i) {
    double *t = &tab[8 * (size_t)i];
    if (i >= 1 && i <= n) { t[0] = -sam[i]; t[1] = -sap[i]; t[2] = se[i]; t[3] = w[i]; t[4] = srw[i]; }
    else t[2] = -1.0;   // outside cell: its update gives 0 and nothing couples to it
  }
  L.tab[d].upload(tab);
}

void MGSolver::setup(Ctx &ctx, std::vector<SBox> &&boxes_, bool batched, int max_levels, int key0_) {
  sboxes = std::move(boxes_);
  fft_dir = batched ? 2 : -1;
  ids_local = batched;
  key0 = key0_;
  distributed = !batched && ctx.comm.nproc > 1;
  if (distributed && key0 < 0) throw Error("multigrid over several processes: the level-0 iterate must be a published array");
  const int nl = (int)sboxes.size();
  bool coupled[3];
  for (int d = 0; d < 3; ++d) coupled[d] = (d != fft_dir);
  nq = batched ? sboxes[0].n[2] : 1;
  norms.alloc(3 * (size_t)nq);
  h_norms.assign(2 * (size_t)nq, 0.0);
  if (!d_active) {
    SNAC_CUDA(cudaMalloc(&d_active, (3 * (size_t)nq + 4) * sizeof(int)));
    d_where = d_active + nq; d_iters = d_where + nq; d_count = d_iters + nq;
    SNAC_CUDA(cudaMemset(d_active, 0, (3 * (size_t)nq + 4) * sizeof(int)));
    SNAC_CUDA(cudaMallocHost(&h_count, (16 + 2 * HIST_KEYS * HIST_LEN) * sizeof(int)));
    for (int i = 0; i < 16 + 2 * HIST_KEYS * HIST_LEN; ++i) h_count[i] = 0;
  }
  lam.clear();
  lam.resize(nl);
  std::vector<NbTable> nbt(nl);
  for (int lb = 0; lb < nl; ++lb) {
    if (batched) lam[lb].upload(sboxes[lb].lam);
    for (int d = 0; d < 3; ++d)
      for (int ib = 0; ib < 2; ++ib) nbt[lb].nb[d][ib] = coupled[d] ? sboxes[lb].nb[d][ib] : -1;
  }

  levels.clear();
  levels.emplace_back(new MGLevel);
  {  // level 0: the reference's own matrix
    MGLevel &L0 = *levels[0];
    L0.boxes.resize(nl);
    x0.clear(); b0.clear();
    x0.resize(nl); b0.resize(nl);
    for (int lb = 0; lb < nl; ++lb) {
      SBox &sb = sboxes[lb];
      MGLevelBox &L = L0.boxes[lb];
      L.g = Grid3::make(sb.n[0], sb.n[1], sb.n[2]);
      if (sb.sy) { L.g.sy = sb.sy; L.g.sz = sb.sz; }
      if (sb.x && sb.b) { L.xp = sb.x; L.bp = sb.b; }
      else {
        x0[lb].alloc(L.g); b0[lb].alloc(L.g);
        L.xp = x0[lb].p; L.bp = b0[lb].p;
      }
      for (int d = 0; d < 3; ++d) build_dir(L, d, coupled[d], sb.d1[d], sb.d2[d], sb.sgn[d], sb.phys[d], sb.qqq[d]);
    }
  }
  // distributed: the layout and the cell widths of EVERY box of the run, level by level (host arithmetic only: each
  // process repeats the coarsening of its neighbours' boxes instead of exchanging ghost widths on the device)
  std::vector<std::array<std::vector<double>, 3>> all_w;   // [rank][d]: widths 1..n of the current level (ghost entries unused)
  if (distributed) {
    const int nr = ctx.prm.nrank;
    rank_grids.assign(1, std::vector<Grid3>(nr));
    all_w.resize(nr);
    for (int r = 0; r < nr; ++r) {
      const snac_box &w = ctx.world[r];
      Grid3 g = Grid3::make(w.hi[0] - w.lo[0] + 1, w.hi[1] - w.lo[1] + 1, w.hi[2] - w.lo[2] + 1);
      for (int d = 0; d < 3; ++d) {
        const bool stag = sboxes[0].qqq[d] != 0;   // the same for every box of a solver
        all_w[r][d] = stag ? ctx.world_dc[d][r] : ctx.world_df[d][r];
        if (stag && w.is_bound[d][1]) g.n[d] -= 1;   // the shortened staggered range keeps the strides of the box
      }
      rank_grids[0][r] = g;
    }
  }
  // coarser levels
  for (;;) {
    MGLevel &F = *levels.back();
    // mean width per direction over the local boxes
    double hsum[3] = {0, 0, 0};
    long ncnt[3] = {0, 0, 0};
    bool all_one[3] = {true, true, true};
    for (int lb = 0; lb < nl; ++lb)
      for (int d = 0; d < 3; ++d) {
        const int n = F.boxes[lb].g.n[d];
        if (n != 1) all_one[d] = false;
        for (int i = 1; i <= n; ++i) hsum[d] += F.boxes[lb].h_w[d][i];
        ncnt[d] += n;
      }
    if (distributed) {   // the same decision on every process: statistics over all boxes of the run
      for (int d = 0; d < 3; ++d) { hsum[d] = 0.0; ncnt[d] = 0; all_one[d] = true; }
      const std::vector<Grid3> &RG = rank_grids.back();
      for (size_t r = 0; r < RG.size(); ++r)
        for (int d = 0; d < 3; ++d) {
          const int n = RG[r].n[d];
          if (n != 1) all_one[d] = false;
          for (int i = 1; i <= n; ++i) hsum[d] += all_w[r][d][i];
          ncnt[d] += n;
        }
    }
    double hmin = 1e300;
    for (int d = 0; d < 3; ++d)
      if (coupled[d] && !all_one[d]) hmin = std::min(hmin, hsum[d] / ncnt[d]);
    int co[3] = {0, 0, 0};
    bool any = false;
    for (int d = 0; d < 3; ++d) {
      if (!coupled[d] || all_one[d]) continue;
      if (hsum[d] / ncnt[d] < 1.5 * hmin) { co[d] = 1; any = true; }
    }
    if (!any || (int)levels.size() >= max_levels) break;
    for (int d = 0; d < 3; ++d) F.coarsen[d] = co[d];
    levels.emplace_back(new MGLevel);
    MGLevel &Fm = *levels[levels.size() - 2];
    MGLevel &C = *levels.back();
    C.boxes.resize(nl);
    for (int lb = 0; lb < nl; ++lb) {
      MGLevelBox &f = Fm.boxes[lb];
      MGLevelBox &c = C.boxes[lb];
      c.g = Grid3::make(co[0] ? (f.g.n[0] + 1) / 2 : f.g.n[0], co[1] ? (f.g.n[1] + 1) / 2 : f.g.n[1],
                        co[2] ? (f.g.n[2] + 1) / 2 : f.g.n[2]);
      if (distributed) c.x.adopt(ctx.comm.alloc(c.g.alloc_elems(), ctx.boxes[lb]->info.rank, AK_MGX + (int)levels.size() - 1));
      else c.x.alloc(c.g);
      c.b.alloc(c.g);
      c.xp = c.x.p; c.bp = c.b.p;
      for (int d = 0; d < 3; ++d) {
        const int n = c.g.n[d];
        c.h_w[d].assign(n + 2, 1.0);
        if (!coupled[d]) continue;
        for (int i = 1; i <= n; ++i)
          c.h_w[d][i] = co[d] ? f.h_w[d][2 * i - 1] + (2 * i <= f.g.n[d] ? f.h_w[d][2 * i] : 0.0) : f.h_w[d][i];
        c.h_w[d][0] = c.h_w[d][1];
        c.h_w[d][n + 1] = c.h_w[d][n];
      }
    }
    if (distributed) {
      // coarse widths and layouts of every box of the run; ghost widths of the local boxes from their neighbours'
      const std::vector<Grid3> &RF = rank_grids[rank_grids.size() - 1];
      std::vector<Grid3> RC(RF.size());
      std::vector<std::array<std::vector<double>, 3>> cw(RF.size());
      for (size_t r = 0; r < RF.size(); ++r) {
        RC[r] = Grid3::make(co[0] ? (RF[r].n[0] + 1) / 2 : RF[r].n[0], co[1] ? (RF[r].n[1] + 1) / 2 : RF[r].n[1],
                            co[2] ? (RF[r].n[2] + 1) / 2 : RF[r].n[2]);
        for (int d = 0; d < 3; ++d) {
          const int n = RC[r].n[d], nf = RF[r].n[d];
          cw[r][d].assign(n + 2, 1.0);
          for (int i = 1; i <= n; ++i)
            cw[r][d][i] = co[d] ? all_w[r][d][2 * i - 1] + (2 * i <= nf ? all_w[r][d][2 * i] : 0.0) : all_w[r][d][i];
        }
      }
      for (int lb = 0; lb < nl; ++lb) {
        MGLevelBox &c = C.boxes[lb];
        for (int d = 0; d < 3; ++d) {
          if (!coupled[d]) continue;
          const int n = c.g.n[d];
          for (int ib = 0; ib < 2; ++ib) {
            const int nbr = nbt[lb].nb[d][ib];
            if (nbr < 0) continue;
            c.h_w[d][ib == 0 ? 0 : n + 1] = ib == 0 ? cw[nbr][d][RC[nbr].n[d]] : cw[nbr][d][1];
          }
        }
      }
      all_w.swap(cw);
      rank_grids.push_back(RC);
      for (int lb = 0; lb < nl; ++lb)
        SNAC_CUDA(cudaMemsetAsync(C.boxes[lb].x.p - 15, 0, C.boxes[lb].g.alloc_elems() * sizeof(double), ctx.stream));
    } else
    // ghost widths across box interfaces: exchange a field that carries the width along d
    {
      std::vector<ExField> ef;
      for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, C.boxes[lb].xp, C.boxes[lb].g});
      C.ex.build(ctx, ef, &nbt, nullptr, ids_local);
      for (int d = 0; d < 3; ++d) {
        if (!coupled[d] || !C.ex.local[d].count) continue;
        std::vector<Dev1> tmp(nl);
        for (int lb = 0; lb < nl; ++lb) {
          MGLevelBox &c = C.boxes[lb];
          tmp[lb].upload(c.h_w[d]);
          dim3 blk(32, 4, 2), grd = mg_grid(blk, c.g.n[0], c.g.n[1], c.g.n[2]);
          mg_fill_along_kernel<<<grd, blk, 0, ctx.stream>>>(c.xp, c.g.sy, c.g.sz, c.g.n[0], c.g.n[1], c.g.n[2], d, tmp[lb].p());
        }
        C.ex.run(ctx, d);
        SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
        for (int lb = 0; lb < nl; ++lb) {
          MGLevelBox &c = C.boxes[lb];
          for (int ib = 0; ib < 2; ++ib) {
            if (nbt[lb].nb[d][ib] < 0) continue;
            long idx[3] = {1, 1, 1};
            idx[d] = ib == 0 ? 0 : c.g.n[d] + 1;
            double v;
            SNAC_CUDA(cudaMemcpy(&v, c.xp + idx[0] + c.g.sy * idx[1] + c.g.sz * idx[2], sizeof(double), cudaMemcpyDeviceToHost));
            c.h_w[d][idx[d]] = v;
          }
        }
      }
      for (int lb = 0; lb < nl; ++lb)
        SNAC_CUDA(cudaMemsetAsync(C.boxes[lb].x.buf.raw, 0, C.boxes[lb].x.buf.elems * sizeof(double), ctx.stream));
    }
    for (int lb = 0; lb < nl; ++lb) {
      SBox &sb = sboxes[lb];
      MGLevelBox &c = C.boxes[lb];
      for (int d = 0; d < 3; ++d) {
        const int n = c.g.n[d];
        std::vector<double> d1(n + 2, 1.0);
        if (coupled[d]) {
          for (int i = 0; i <= n; ++i) d1[i] = 0.5 * (c.h_w[d][i] + c.h_w[d][i + 1]);
          d1[n + 1] = d1[n];
        }
        build_dir(c, d, coupled[d], d1, c.h_w[d], sb.sgn[d], sb.phys[d]);
      }
    }
  }
  // halo exchange of x on level 0 (coarser ones were built above); several processes: once the peers are mapped
  if (!distributed) {
    std::vector<ExField> ef;
    for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, levels[0]->boxes[lb].xp, levels[0]->boxes[lb].g});
    levels[0]->ex.build(ctx, ef, &nbt, nullptr, ids_local);
  }
  // ---- plane-local fast path -----------------------------------------------------------------
  plane_local = batched && !ctx.prm.mg_generic;
  int npre, npost;
  mg_sweeps(ctx, *this, npre, npost);
  if (npre != npost || npre > 2) plane_local = false;
  for (int lb = 0; lb < nl && plane_local; ++lb)
    for (int d = 0; d < 2; ++d)
      for (int ib = 0; ib < 2; ++ib) {
        const int nb = sboxes[lb].nb[d][ib];
        if (nb >= 0 && (ids_local ? nb : ctx.local_of_rank[nb]) != lb) plane_local = false;   // the plane is split over boxes
      }
  first_smem_level = 0;
  // test hook: the largest plane handed to the one-CTA-per-plane kernel (forces the streaming kernels on small cases)
  const size_t smem_cells = getenv("SNAC_MG_SMEM_CELLS") ? (size_t)atol(getenv("SNAC_MG_SMEM_CELLS")) : 0;
  if (plane_local) {
    const int nlev = (int)levels.size();
    // first level whose whole hierarchy below fits in shared memory (x and b per level): in FP64
    // when that is level 0 (the kernel is then the whole solver), in FP32 below it (mg_coarse.cu)
    int ls = nlev - 1;
    coarse_single = false;
    for (int l = nlev - 1; l >= 0; --l) {
      bool ok = true;
      size_t worst = 0, tabs = 0;
      for (int lb = 0; lb < nl; ++lb) {
        size_t c = 0, t = 0;
        for (int m = l; m < nlev; ++m) {
          c += (size_t)levels[m]->boxes[lb].g.n[0] * levels[m]->boxes[lb].g.n[1];
          t += (size_t)levels[m]->boxes[lb].g.n[0] + levels[m]->boxes[lb].g.n[1];
        }
        worst = std::max(worst, c);
        tabs = std::max(tabs, t);
      }
      size_t top = 0;
      for (int lb = 0; lb < nl; ++lb) top = std::max(top, (size_t)levels[l]->boxes[lb].g.n[0] * levels[l]->boxes[lb].g.n[1]);
      const size_t elems = 2 * worst + 5 * tabs;
      const size_t bytes = elems * (l == 0 ? sizeof(double) : sizeof(float));
      if (nlev - l > CMAXLEV || bytes > 200 * 1024) ok = false;
      if (smem_cells && l < nlev - 1 && top > smem_cells) ok = false;
      if (!ok) break;
      ls = l;
    }
    coarse_single = ls > 0;
    first_smem_level = ls;
    // the tiled kernels need an even size in every periodic direction of the levels they handle
    for (int l = 0; l < ls && plane_local; ++l)
      for (int lb = 0; lb < nl; ++lb)
        for (int d = 0; d < 2; ++d)
          if (sboxes[lb].nb[d][0] >= 0 && (levels[l]->boxes[lb].g.n[d] % 2)) plane_local = false;
    if (plane_local) {
      x1.clear();
      x1.resize((size_t)ls * nl);
      for (int l = 0; l < ls; ++l)
        for (int lb = 0; lb < nl; ++lb) {
          if (l == 0 && sboxes[lb].xalt[0]) x1[lb].p = sboxes[lb].xalt[0];   // arena memory (several processes)
          else x1[(size_t)l * nl + lb].alloc(levels[l]->boxes[lb].g);
        }
      xslot.assign(nl, {nullptr, nullptr, nullptr, nullptr});
      for (int lb = 0; lb < nl; ++lb) {
        xslot[lb][0] = levels[0]->boxes[lb].xp;
        xslot[lb][1] = ls > 0 ? x1[lb].p : nullptr;
        xslot[lb][2] = sboxes[lb].xalt[1];
        xslot[lb][3] = sboxes[lb].xalt[2];
      }
      slot_primary = 0; slot_scratch = 1; cur_key = 0;
      slot_key[0] = -1; slot_key[1] = 2; slot_key[2] = 3; slot_key[3] = -1;
    }
  }
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
}

// several processes: halo tables of every level, neighbours' arrays addressed through the mapped arenas
void MGSolver::connect(Ctx &ctx) {
  if (!distributed) return;
  const int nl = (int)sboxes.size();
  std::vector<NbTable> nbt(nl);
  for (int lb = 0; lb < nl; ++lb)
    for (int d = 0; d < 3; ++d)
      for (int ib = 0; ib < 2; ++ib) nbt[lb].nb[d][ib] = (d != fft_dir) ? sboxes[lb].nb[d][ib] : -1;
  for (size_t l = 0; l < levels.size(); ++l) {
    std::vector<ExField> ef;
    for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, levels[l]->boxes[lb].xp, levels[l]->boxes[lb].g});
    const std::vector<int> keys = {l == 0 ? key0 : AK_MGX + (int)l};
    levels[l]->ex.build(ctx, ef, &nbt, &keys, false, &rank_grids[l], false);
  }
}

namespace {
struct Cycle {
  Ctx &ctx;
  MGSolver &S;
  bool coupled[3];
  int npre, npost, nbot;

  void exchange(int l) {
    ProfScope ps(ctx, l == 0 ? PC_HALO_BC : PC_MG_COARSE);
    for (int d = 0; d < 3; ++d)
      if (coupled[d]) S.levels[l]->ex.run(ctx, d);
  }
  void smooth(int l, int color) {
    ProfScope ps(ctx, l == 0 ? PC_MG_SMOOTH0 : PC_MG_COARSE);
    MGLevel &L = *S.levels[l];
    for (size_t lb = 0; lb < L.boxes.size(); ++lb) {
      MGKernelBox B = kbox(S, L.boxes[lb], (int)lb, coupled);
      dim3 blk = mg_block(L.boxes[lb].g, S.fft_dir), grd = mg_grid(blk, B.n[0], B.n[1], B.n[2]);
      mg_smooth_kernel<<<grd, blk, 0, ctx.stream>>>(B, color, S.alpha);
      ctx.launches++;
    }
  }
  void sweeps(int l, int n, bool reverse) {
    for (int s = 0; s < n; ++s)
      for (int c = 0; c < 2; ++c) {
        smooth(l, reverse ? 1 - c : c);
        exchange(l);
      }
  }
  void restrict_from(int l, bool norms, bool want_b) {
    ProfScope ps(ctx, l == 0 ? PC_MG_DOWN0 : PC_MG_COARSE);
    MGLevel &F = *S.levels[l];
    const bool has_coarse = l + 1 < (int)S.levels.size();
    for (size_t lb = 0; lb < F.boxes.size(); ++lb) {
      MGKernelBox B = kbox(S, F.boxes[lb], (int)lb, coupled);
      int co[3] = {0, 0, 0};
      Grid3 cg = F.boxes[lb].g;
      double *cb = nullptr;
      if (has_coarse) {
        MGLevelBox &c = S.levels[l + 1]->boxes[lb];
        for (int d = 0; d < 3; ++d) co[d] = F.coarsen[d];
        cg = c.g;
        cb = c.bp;
        SNAC_CUDA(cudaMemsetAsync(c.x.p - 15, 0, c.g.alloc_elems() * sizeof(double), ctx.stream));   // (owned or in the arena)
      }
      dim3 blk = mg_block(cg, S.fft_dir), grd = mg_grid(blk, cg.n[0], cg.n[1], cg.n[2]);
      mg_restrict_kernel<<<grd, blk, 0, ctx.stream>>>(B, cb, cg.sy, cg.sz, cg.n[0], cg.n[1], cg.n[2], co[0], co[1], co[2],
                                                      S.alpha, norms ? S.norms.raw : nullptr, S.nq, want_b ? 1 : 0);
      ctx.launches++;
    }
  }
  void prolong_to(int l) {
    ProfScope ps(ctx, l == 0 ? PC_MG_UP0 : PC_MG_COARSE);
    MGLevel &F = *S.levels[l];
    for (size_t lb = 0; lb < F.boxes.size(); ++lb) {
      MGKernelBox B = kbox(S, F.boxes[lb], (int)lb, coupled);
      MGLevelBox &c = S.levels[l + 1]->boxes[lb];
      dim3 blk = mg_block(F.boxes[lb].g, S.fft_dir), grd = mg_grid(blk, B.n[0], B.n[1], B.n[2]);
      mg_prolong_kernel<<<grd, blk, 0, ctx.stream>>>(B, c.xp, c.g.sy, c.g.sz, F.coarsen[0], F.coarsen[1], F.coarsen[2]);
      ctx.launches++;
    }
  }
  // coarse-grid correction starting at level l (x on level l is zero on entry for l > 0)
  void descend(int l) {
    const int nlev = (int)S.levels.size();
    if (l == nlev - 1) {
      sweeps(l, nbot, false);
      return;
    }
    sweeps(l, npre, false);
    restrict_from(l, false, false);
    descend(l + 1);
    prolong_to(l);
    exchange(l);
    sweeps(l, npost, true);
  }

  // ---- plane-local fast path ---------------------------------------------------------------
  RLevel rlevel(int l, int lb) const {
    const MGLevelBox &B = S.levels[l]->boxes[lb];
    RLevel R;
    R.n1 = B.g.n[0]; R.n2 = B.g.n[1];
    R.sy = (int)B.g.sy; R.sz = B.g.sz;
    R.ctab = B.tab[0].p(); R.rtab = B.tab[1].p();
    R.c1 = S.levels[l]->coarsen[0]; R.c2 = S.levels[l]->coarsen[1];
    R.per1 = S.sboxes[lb].nb[0][0] >= 0; R.per2 = S.sboxes[lb].nb[1][0] >= 0;
    return R;
  }
  double *xbuf(int l, int lb, int which) const {   // ping-pong buffers of the tiled levels
    return which == 0 ? S.levels[l]->boxes[lb].xp : S.x1[(size_t)l * S.levels[0]->boxes.size() + lb].p;
  }
  // down: xbuf(l,0) -> xbuf(l,1), residual restricted into level l+1's b
  double active_points(int l) {   // profile mode only: planes still iterating x plane size
    if (!ctx.prof_on) return 0.0;
    int cnt = S.nq;
    cudaMemcpyAsync(&cnt, S.d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream);
    cudaStreamSynchronize(ctx.stream);
    double cells = 0.0;
    for (auto &b : S.levels[l]->boxes) cells += (double)b.g.n[0] * b.g.n[1];
    return cells * cnt;
  }
  void plane_down(int l, bool norms, bool want_b) {
    ProfScope ps(ctx, l == 0 ? PC_MG_DOWN0 : PC_MG_COARSE, l == 0 ? active_points(0) : 0.0);
    MGLevel &F = *S.levels[l];
    for (size_t lb = 0; lb < F.boxes.size(); ++lb) {
      MGLevelBox &c = S.levels[l + 1]->boxes[lb];
      const double *xi = l == 0 ? xbuf(l, (int)lb, 0) : nullptr;   // below level 0 the iterate is zero on entry
      launch_ring_down(ctx.stream, rlevel(l, (int)lb), npre, xi, xbuf(l, (int)lb, 1), F.boxes[lb].bp, c.bp, c.g.sy, c.g.sz, S.lam[lb].p(),
                       S.alpha, norms ? S.norms.raw : nullptr, S.nq, want_b ? 1 : 0, S.d_active, S.est_active, (l == 0 && want_b) ? S.guess_scale : 1.0);
      ctx.launches++;
    }
  }
  void plane_up(int l) {
    ProfScope ps(ctx, l == 0 ? PC_MG_UP0 : PC_MG_COARSE, l == 0 ? active_points(0) : 0.0);
    MGLevel &F = *S.levels[l];
    for (size_t lb = 0; lb < F.boxes.size(); ++lb) {
      MGLevelBox &c = S.levels[l + 1]->boxes[lb];
      launch_ring_up(ctx.stream, rlevel(l, (int)lb), npost, xbuf(l, (int)lb, 1), xbuf(l, (int)lb, 0), F.boxes[lb].bp, c.xp, c.g.sy, c.g.sz,
                     S.lam[lb].p(), S.alpha, S.nq, S.d_active, S.est_active);
      ctx.launches++;
    }
  }
  void plane_coarse(bool norms, bool want_b) {
    const int ls = S.first_smem_level, nlev = (int)S.levels.size();
    ProfScope ps(ctx, ls == 0 ? PC_MG_DOWN0 : PC_MG_COARSE);
    for (size_t lb = 0; lb < S.levels[ls]->boxes.size(); ++lb) {
      CoarseArgs a;
      a.nlev = nlev - ls;
      for (int l = ls; l < nlev; ++l) {
        const MGLevelBox &B = S.levels[l]->boxes[lb];
        CoarseLevelArg &c = a.lev[l - ls];
        c.n1 = B.g.n[0]; c.n2 = B.g.n[1];
        c.c1 = S.levels[l]->coarsen[0]; c.c2 = S.levels[l]->coarsen[1];
        c.per1 = S.sboxes[lb].nb[0][0] >= 0; c.per2 = S.sboxes[lb].nb[1][0] >= 0;
        c.t1 = B.tab[0].p(); c.t2 = B.tab[1].p();
      }
      MGLevelBox &T = S.levels[ls]->boxes[lb];
      a.b = T.bp;
      a.xin = ls == 0 ? T.xp : nullptr;
      a.xout = T.xp;
      a.sy = T.g.sy; a.sz = T.g.sz;
      a.lamq = S.lam[lb].p();
      a.alpha = S.alpha;
      a.norms = (ls == 0 && norms) ? S.norms.raw : nullptr;
      a.xscale = (ls == 0 && want_b) ? S.guess_scale : 1.0;
      a.nq = S.nq; a.want_b = want_b ? 1 : 0;
      a.npre = npre; a.npost = npost; a.nbot = nbot;
      a.npre_low = npre; a.npost_low = npost;
      a.adaptive = ls > 0 ? 1 : 0;
      a.active = S.d_active;
      launch_mg_coarse(ctx.stream, a, S.coarse_single);
      ctx.launches++;
    }
  }
};
}  // namespace

void MGSolver::stats(Ctx &ctx, int *iters, double *relres, double *mean_iters) {
  mg_coop_collect(ctx, *this, coop);
  if (plane_local && !direct.on) {
    std::vector<int> it(nq);
    std::vector<double> rel(nq);
    SNAC_CUDA(cudaMemcpyAsync(it.data(), d_iters, nq * sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    SNAC_CUDA(cudaMemcpyAsync(rel.data(), norms.raw + 2 * (size_t)nq, nq * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
    last_iters = 0; last_relres = 0.0; last_mean_iters = 0.0;
    for (int q = 0; q < nq; ++q) {
      last_iters = std::max(last_iters, it[q]);
      last_relres = std::max(last_relres, rel[q]);
      last_mean_iters += it[q];
    }
    last_mean_iters /= nq;
  }
  if (iters) *iters = last_iters;
  if (relres) *relres = last_relres;
  if (mean_iters) *mean_iters = last_mean_iters;
}

// plane-local path: the stopping test runs on the device, per wavenumber; the host only polls the
// number of wavenumbers still iterating once per batch of cycles
static void solve_plane_local(Ctx &ctx, MGSolver &S, Cycle &cy) {
  const int nq = S.nq, ls = S.first_smem_level;
  const double tol = ctx.prm.hypre_tol;
  const int maxiter = std::max(ctx.prm.hypre_maxiter, 1);
  const bool in_alt = ls > 0;
  SNAC_CUDA(cudaMemsetAsync(S.norms.raw, 0, 3 * (size_t)nq * sizeof(double), ctx.stream));
  mg_flags_init_kernel<<<1, 256, 0, ctx.stream>>>(S.d_active, S.d_where, S.d_iters, S.d_count, nq);
  ctx.launches++;
  int it = 0;
  // record of the last solve with the same key (see MGSolver::hist)
  const int key = S.hist_key & (MGSolver::HIST_KEYS - 1);
  const int *prev = S.hist(key, 1 - S.hist_cur[key]);
  int *cur = S.hist(key, S.hist_cur[key]);
  const int prev_n = S.hist_n[key];
  int polled = nq;
  auto predict = [&](int i) -> int {   // planes still iterating after cycle i
    if (i < 0) return nq;
    if (i < prev_n) return std::min(nq, std::max(1, prev[i] + prev[i] / 8 + 1));
    return std::max(1, polled);
  };
  {   // initial guess: the last solution of this key (see MGSolver::xslot)
    if (in_alt && key != S.cur_key && S.slot_key[key] >= 0) {
      MGLevel &L0 = *S.levels[0];
      const int snew = S.slot_key[key];
      for (size_t lb = 0; lb < L0.boxes.size(); ++lb) {
        if (!S.xslot[lb][snew]) {   // one process, first solve with this key: its own array, zero guess
          if (ctx.comm.nproc > 1) throw Error("solver: the per-key iterate arrays were not placed in the arena");
          if (S.slot_own[snew].size() != L0.boxes.size()) S.slot_own[snew].resize(L0.boxes.size());
          S.slot_own[snew][lb].alloc(L0.boxes[lb].g);
          S.xslot[lb][snew] = S.slot_own[snew][lb].p;
        }
        L0.boxes[lb].xp = S.xslot[lb][snew];
      }
      S.slot_key[S.cur_key] = S.slot_primary;   // the current array goes to rest with the solution of its key
      S.slot_primary = snew;
      S.slot_key[key] = -1;
      S.cur_key = key;
      S.swapped = true;
    }
  }
  bool first = true;
  for (;;) {
    // cycles enqueued before the host looks at the counter: as many as the last solve needed, then two at a time
    const int batch = first ? (prev_n > 0 ? std::min(prev_n, maxiter + 1) : 3) : 2;
    first = false;
    for (int c = 0; c < batch; ++c, ++it) {
      S.est_active = predict(it - 1);
      if (ls == 0) cy.plane_coarse(true, it == 0);   // the whole cycle in one kernel
      else cy.plane_down(0, true, it == 0);
      {
        ProfScope ps(ctx, PC_MG_COARSE);
        mg_flags_kernel<<<1, 256, 0, ctx.stream>>>(S.norms.raw, nq, tol * tol, it, maxiter, in_alt ? 1 : 0, S.d_active, S.d_where,
                                                   S.d_iters, S.d_count);
        ctx.launches++;
      }
      if (it < MGSolver::HIST_LEN)
        SNAC_CUDA(cudaMemcpyAsync(cur + it, S.d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
      if (ls > 0) {
        S.est_active = predict(it);
        for (int l = 1; l < ls; ++l) cy.plane_down(l, false, false);
        cy.plane_coarse(false, false);
        for (int l = ls - 1; l >= 0; --l) cy.plane_up(l);
      }
    }
    SNAC_CUDA(cudaMemcpyAsync(S.h_count, S.d_count, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
    if (*S.h_count == 0 || it > maxiter) break;
    polled = *S.h_count;
  }
  {   // the record of this solve becomes the prediction of the next one with the same key
    int n = 0;
    const int ran = std::min(it, (int)MGSolver::HIST_LEN);
    while (n < ran && cur[n] != 0) ++n;
    S.hist_n[key] = n < ran ? n + 1 : 0;   // cycles until the counter reached zero; no record if it never did
    S.hist_cur[key] = 1 - S.hist_cur[key];
  }
  {
    ProfScope ps(ctx, PC_MG_COARSE);
    MGLevel &L = *S.levels[0];
    // With the ping-pong buffers every converged plane ends in the second buffer (the stopping test
    // follows a down kernel): the buffers trade places and only planes with a zero right-hand side are
    // written (x = 0), instead of a copy of the whole solution (16 B per point per solve); the caller
    // re-points the transform at the new array, its own and its peers' (MGSolver::swapped, xslot).
    const bool swap = in_alt;
    if (swap) std::swap(S.slot_primary, S.slot_scratch);
    for (size_t lb = 0; lb < L.boxes.size(); ++lb) {
      const Grid3 &g = L.boxes[lb].g;
      dim3 blk(64, 4, 1), grd((g.n[0] + 63) / 64, (g.n[1] + 3) / 4, nq);
      if (swap) { L.boxes[lb].xp = S.xslot[lb][S.slot_primary]; S.x1[lb].p = S.xslot[lb][S.slot_scratch]; }
      mg_finalize_kernel<<<grd, blk, 0, ctx.stream>>>(cy.xbuf(0, (int)lb, 0), in_alt ? cy.xbuf(0, (int)lb, 1) : cy.xbuf(0, (int)lb, 0),
                                                      g.sy, g.sz, g.n[0], g.n[1], S.d_where, swap ? 1 : 0);
      ctx.launches++;
    }
    if (swap) S.swapped = true;
  }
}

void MGSolver::solve(Ctx &ctx) {
  if (direct.on) { direct.solve(ctx, *this); return; }
  int npre, npost;
  mg_sweeps(ctx, *this, npre, npost);
  Cycle cy{ctx, *this, {fft_dir != 0, fft_dir != 1, fft_dir != 2}, npre, npost,
           ctx.prm.mg_nbottom > 0 ? ctx.prm.mg_nbottom : 20};
  if (plane_local) {
    solve_plane_local(ctx, *this, cy);
    return;
  }
  // general multi-box path.  Every coupled box on this process: one cooperative launch per solve (mg_coop.cu)
  if (!coop_tried) { coop = mg_coop_setup(ctx, *this, cy.coupled); coop_tried = true; }
  if (coop) {
    ProfScope ps(ctx, PC_MG_SMOOTH0);
    mg_coop_solve(ctx, *this, coop, cy.npre, cy.npost, cy.nbot);
    return;
  }
  // boxes coupled across processes: one launch per box, colour and level; the host reads the residual norms once per cycle
  const int nlev = (int)levels.size();
  const double tol = ctx.prm.hypre_tol;
  const int maxiter = std::max(ctx.prm.hypre_maxiter, 1);
  last_iters = 0;
  last_relres = 0.0;
  cy.exchange(0);
  for (int it = 0;; ++it) {
    const size_t nn = (it == 0 ? 2 : 1) * (size_t)nq * sizeof(double);
    SNAC_CUDA(cudaMemsetAsync(norms.raw, 0, nn, ctx.stream));
    // pre-relaxation on level 0 + residual (+ restriction); its norm drives the stopping test
    if (nlev == 1) cy.sweeps(0, cy.nbot, false);
    else cy.sweeps(0, cy.npre, false);
    cy.restrict_from(0, true, it == 0);
    if (distributed) {
      // MPI_ALLREDUCE of the residual norms (HYPRE's stopping test is global): every process gets the same sums and
      // takes the same decisions below; the sums go back to the device for mg_zero_planes_kernel
      if (nq != 1) throw Error("multigrid over several processes: one system per solve expected");
      double sum[3], mx[3];
      ctx.peer_reduce(norms.raw, 2, sum, mx);
      h_norms[0] = sum[0];
      if (it == 0) h_norms[1] = sum[1];
      SNAC_CUDA(cudaMemcpyAsync(norms.raw, h_norms.data(), nn, cudaMemcpyHostToDevice, ctx.stream));
    } else {
      SNAC_CUDA(cudaMemcpyAsync(h_norms.data(), norms.raw, nn, cudaMemcpyDeviceToHost, ctx.stream));
      SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
    }
    bool conv = true, any_zero = false;
    double worst = 0.0;
    for (int q = 0; q < nq; ++q) {
      const double b2 = h_norms[nq + q], r2 = h_norms[q];
      if (b2 == 0.0) { any_zero = true; continue; }
      const double rel = std::sqrt(r2 / b2);
      worst = std::max(worst, rel);
      if (!(rel < tol)) conv = false;
    }
    last_relres = worst;
    if (it == 0 && any_zero) {   // HYPRE returns x = 0 for b = 0
      MGLevel &L = *levels[0];
      for (size_t lb = 0; lb < L.boxes.size(); ++lb) {
        MGKernelBox B = kbox(*this, L.boxes[lb], (int)lb, cy.coupled);
        dim3 blk = mg_block(L.boxes[lb].g, fft_dir), grd = mg_grid(blk, B.n[0], B.n[1], B.n[2]);
        mg_zero_planes_kernel<<<grd, blk, 0, ctx.stream>>>(B, norms.raw, nq);
        ctx.launches++;
      }
      cy.exchange(0);
      if (worst > 0.0) cy.restrict_from(0, false, false);   // the coarse RHS used the old iterate there
    }
    const bool stop = (it > 0 && conv) || it >= maxiter || (worst == 0.0) || (nlev == 1 && conv);
    if (stop) {
      last_iters = it;
      last_mean_iters = it;
      break;
    }
    if (nlev > 1) {
      cy.descend(1);
      cy.prolong_to(0);
      cy.exchange(0);
      cy.sweeps(0, cy.npost, true);
    }
  }
}

}  // namespace snac
