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
// grids), and only while every box keeps an even size.
#include <algorithm>
#include <cmath>

#include "common.h"

namespace snac {

// ------------------------------------------------------------------------------------------
// device kernels
// ------------------------------------------------------------------------------------------
struct MGCoef {  // coefficients of one cell
  double diag, w[3];
};

__device__ inline double mg_offdiag(const MGKernelBox &B, const double *x, long c, int i, int j, int k,
                                    double wx, double wy, double wz) {
  double off = 0.0;
  if (B.coupled[0]) off += wy * wz * (__ldg(B.am[0] + i) * x[c - 1] + __ldg(B.ap[0] + i) * x[c + 1]);
  if (B.coupled[1]) off += wx * wz * (__ldg(B.am[1] + j) * x[c - B.sy] + __ldg(B.ap[1] + j) * x[c + B.sy]);
  if (B.coupled[2]) off += wx * wy * (__ldg(B.am[2] + k) * x[c - B.sz] + __ldg(B.ap[2] + k) * x[c + B.sz]);
  return off;
}

__device__ inline double mg_diag(const MGKernelBox &B, int i, int j, int k, double wx, double wy, double wz,
                                 double alpha) {
  double lam = alpha;
  if (B.fft_dir >= 0) lam += __ldg(B.lam + (B.fft_dir == 0 ? i : (B.fft_dir == 1 ? j : k)));
  return wy * wz * __ldg(B.ad[0] + i) + wx * wz * __ldg(B.ad[1] + j) + wx * wy * __ldg(B.ad[2] + k) + lam * wx * wy * wz;
}

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
    const int i0 = c0 ? 2 * I - 1 : I, i1 = c0 ? 2 * I : I;
    const int j0 = c1 ? 2 * J - 1 : J, j1 = c1 ? 2 * J : J;
    const int k0 = c2 ? 2 * K - 1 : K, k1 = c2 ? 2 * K : K;
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
// host
// ------------------------------------------------------------------------------------------
static dim3 mg_block(const Grid3 &g, int fft_dir) {
  // threads of a block must share the plane index along y or z when that is the batch direction
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
                      const double sgn[2], const bool phys[2]) {
  const int n = L.g.n[d];
  std::vector<double> am(n + 2, 0.0), ap(n + 2, 0.0), ad(n + 2, 0.0), w(n + 2, 1.0);
  if (coupled) {
    for (int i = 0; i <= n + 1; ++i) w[i] = d2[i];
    for (int i = 1; i <= n; ++i) {
      am[i] = 1.0 / d1[i - 1];
      ap[i] = 1.0 / d1[i];
      ad[i] = -(am[i] + ap[i]);
    }
    if (phys[0]) { ad[1] += sgn[0] * am[1]; am[1] = 0.0; }
    if (phys[1]) { ad[n] += sgn[1] * ap[n]; ap[n] = 0.0; }
  }
  L.h_w[d] = w;
  L.am[d].upload(am); L.ap[d].upload(ap); L.ad[d].upload(ad); L.w[d].upload(w);
}

void MGSolver::setup(Ctx &ctx, const bool centered_[3]) {
  for (int d = 0; d < 3; ++d) centered[d] = centered_[d];
  fft_dir = ctx.prm.fft_axis - 1;
  const int nl = (int)ctx.boxes.size();
  bool coupled[3];
  for (int d = 0; d < 3; ++d) coupled[d] = (d != fft_dir);
  nq = fft_dir >= 0 ? ctx.boxes[0]->g.n[fft_dir] : 1;
  norms.alloc(2 * (size_t)nq);
  h_norms.assign(2 * (size_t)nq, 0.0);

  // eigenvalues of the transform direction (src/fft.f90:151-189, src/solver.f90:1124-1138)
  lam.clear();
  lam.resize(nl);
  if (fft_dir >= 0) {
    for (int lb = 0; lb < nl; ++lb) {
      Box &bx = *ctx.boxes[lb];
      const char bc[2] = {bx.info.cbcpre[fft_dir][0], bx.info.cbcpre[fft_dir][1]};
      std::vector<double> l = r2r_eigenvalues(bx.g.n[fft_dir], bc, true);
      const double dl = bx.info.dl[fft_dir][0];
      std::vector<double> lg(l.size() + 2, 0.0);
      for (size_t q = 0; q < l.size(); ++q) lg[q + 1] = l[q] / (dl * dl);
      lam[lb].upload(lg);
    }
  }

  // sign of the boundary fold per face of each box (src/solver.f90:71-94)
  auto face_sgn = [&](const Box &bx, int d, int ib, bool &phys) {
    phys = bx.info.is_bound[d][ib] && ctx.prm.periods[d] == 0;
    if (!phys) return 0.0;
    const char c = bx.info.cbcpre[d][ib];
    if (c == 'N') return 1.0;
    if (c == 'D') return centered[d] ? -1.0 : 0.0;
    throw Error("multigrid set-up: a physical boundary must be 'D' or 'N'");
  };

  levels.clear();
  levels.emplace_back(new MGLevel);
  {  // level 0: the reference's own matrix
    MGLevel &L0 = *levels[0];
    L0.boxes.resize(nl);
    for (int lb = 0; lb < nl; ++lb) {
      Box &bx = *ctx.boxes[lb];
      MGLevelBox &L = L0.boxes[lb];
      L.g = bx.g;
      L.xp = bx.po.p;
      L.bp = bx.rhs.p;
      for (int d = 0; d < 3; ++d) {
        bool phys[2];
        double sgn[2];
        for (int ib = 0; ib < 2; ++ib) sgn[ib] = face_sgn(bx, d, ib, phys[ib]);
        build_dir(L, d, coupled[d], bx.h_dc[d], bx.h_df[d], sgn, phys);
      }
    }
  }
  // coarser levels
  for (;;) {
    MGLevel &F = *levels.back();
    // mean width and parity per direction over the local boxes
    double hsum[3] = {0, 0, 0};
    long ncnt[3] = {0, 0, 0};
    bool even[3] = {true, true, true}, all_one[3] = {true, true, true};
    for (int lb = 0; lb < nl; ++lb)
      for (int d = 0; d < 3; ++d) {
        const int n = F.boxes[lb].g.n[d];
        if (n % 2) even[d] = false;
        if (n != 1) all_one[d] = false;
        for (int i = 1; i <= n; ++i) hsum[d] += F.boxes[lb].h_w[d][i];
        ncnt[d] += n;
      }
    double hmin = 1e300;
    for (int d = 0; d < 3; ++d)
      if (coupled[d] && !all_one[d]) hmin = std::min(hmin, hsum[d] / ncnt[d]);
    int co[3] = {0, 0, 0};
    bool any = false;
    for (int d = 0; d < 3; ++d) {
      if (!coupled[d] || !even[d]) continue;
      if (hsum[d] / ncnt[d] < 1.5 * hmin) { co[d] = 1; any = true; }
    }
    if (!any || levels.size() >= 24) break;
    for (int d = 0; d < 3; ++d) F.coarsen[d] = co[d];
    levels.emplace_back(new MGLevel);
    MGLevel &Fm = *levels[levels.size() - 2];
    MGLevel &C = *levels.back();
    C.boxes.resize(nl);
    for (int lb = 0; lb < nl; ++lb) {
      MGLevelBox &f = Fm.boxes[lb];
      MGLevelBox &c = C.boxes[lb];
      c.g = Grid3::make(co[0] ? f.g.n[0] / 2 : f.g.n[0], co[1] ? f.g.n[1] / 2 : f.g.n[1], co[2] ? f.g.n[2] / 2 : f.g.n[2]);
      c.x.alloc(c.g); c.b.alloc(c.g);
      c.xp = c.x.p; c.bp = c.b.p;
      for (int d = 0; d < 3; ++d) {
        const int n = c.g.n[d];
        c.h_w[d].assign(n + 2, 1.0);
        if (!coupled[d]) continue;
        for (int i = 1; i <= n; ++i) c.h_w[d][i] = co[d] ? f.h_w[d][2 * i - 1] + f.h_w[d][2 * i] : f.h_w[d][i];
        c.h_w[d][0] = c.h_w[d][1];
        c.h_w[d][n + 1] = c.h_w[d][n];
      }
    }
    // ghost widths across box interfaces: exchange a field that carries the width along d
    {
      std::vector<ExField> ef;
      for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, C.boxes[lb].xp, C.boxes[lb].g});
      C.ex.build(ctx, ef);
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
          const snac_box &info = ctx.boxes[lb]->info;
          for (int ib = 0; ib < 2; ++ib) {
            if (info.nb[d][ib] < 0) continue;
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
      Box &bx = *ctx.boxes[lb];
      MGLevelBox &c = C.boxes[lb];
      for (int d = 0; d < 3; ++d) {
        const int n = c.g.n[d];
        std::vector<double> d1(n + 2, 1.0);
        if (coupled[d]) {
          if (centered[d]) {
            for (int i = 0; i <= n; ++i) d1[i] = 0.5 * (c.h_w[d][i] + c.h_w[d][i + 1]);
            d1[n + 1] = d1[n];
          } else {
            throw Error("multigrid: staggered operators are not available in this build");
          }
        }
        bool phys[2];
        double sgn[2];
        for (int ib = 0; ib < 2; ++ib) sgn[ib] = face_sgn(bx, d, ib, phys[ib]);
        build_dir(c, d, coupled[d], d1, c.h_w[d], sgn, phys);
      }
    }
  }
  // halo exchange of x on level 0 (coarser ones were built above)
  {
    std::vector<ExField> ef;
    for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, levels[0]->boxes[lb].xp, levels[0]->boxes[lb].g});
    levels[0]->ex.build(ctx, ef);
  }
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
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
    ProfScope ps(ctx, l == 0 ? PC_MG_RESTRICT0 : PC_MG_COARSE);
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
        SNAC_CUDA(cudaMemsetAsync(c.x.buf.raw, 0, c.x.buf.elems * sizeof(double), ctx.stream));
      }
      dim3 blk = mg_block(cg, S.fft_dir), grd = mg_grid(blk, cg.n[0], cg.n[1], cg.n[2]);
      mg_restrict_kernel<<<grd, blk, 0, ctx.stream>>>(B, cb, cg.sy, cg.sz, cg.n[0], cg.n[1], cg.n[2], co[0], co[1], co[2],
                                                      S.alpha, norms ? S.norms.raw : nullptr, S.nq, want_b ? 1 : 0);
      ctx.launches++;
    }
  }
  void prolong_to(int l) {
    ProfScope ps(ctx, l == 0 ? PC_MG_PROLONG0 : PC_MG_COARSE);
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
};
}  // namespace

void MGSolver::solve(Ctx &ctx) {
  Cycle cy{ctx, *this, {fft_dir != 0, fft_dir != 1, fft_dir != 2},
           ctx.prm.mg_npre > 0 ? ctx.prm.mg_npre : 1, ctx.prm.mg_npost > 0 ? ctx.prm.mg_npost : 1,
           ctx.prm.mg_nbottom > 0 ? ctx.prm.mg_nbottom : 20};
  const int nlev = (int)levels.size();
  const double tol = ctx.prm.hypre_tol;
  const int maxiter = std::max(ctx.prm.hypre_maxiter, 1);
  last_iters = 0;
  last_relres = 0.0;
  cy.exchange(0);
  for (int it = 0;; ++it) {
    if (nlev == 1) cy.sweeps(0, cy.nbot, false);
    else cy.sweeps(0, cy.npre, false);
    SNAC_CUDA(cudaMemsetAsync(norms.raw, 0, (it == 0 ? 2 : 1) * (size_t)nq * sizeof(double), ctx.stream));
    cy.restrict_from(0, true, it == 0);
    SNAC_CUDA(cudaMemcpyAsync(h_norms.data(), norms.raw, (it == 0 ? 2 : 1) * (size_t)nq * sizeof(double),
                              cudaMemcpyDeviceToHost, ctx.stream));
    SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
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
    if (it == 0 && any_zero) {
      MGLevel &L = *levels[0];
      for (size_t lb = 0; lb < L.boxes.size(); ++lb) {
        MGKernelBox B = kbox(*this, L.boxes[lb], (int)lb, cy.coupled);
        dim3 blk = mg_block(L.boxes[lb].g, fft_dir), grd = mg_grid(blk, B.n[0], B.n[1], B.n[2]);
        mg_zero_planes_kernel<<<grd, blk, 0, ctx.stream>>>(B, norms.raw, nq);
        ctx.launches++;
      }
      cy.exchange(0);
      if (worst == 0.0) {  // every plane has a zero right-hand side
        last_iters = 0;
        break;
      }
      // the coarse right-hand side was computed with the old iterate in those planes: redo it
      cy.restrict_from(0, false, false);
    }
    if ((it > 0 && conv) || it >= maxiter || nlev == 1 && conv) {
      last_iters = it;
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
