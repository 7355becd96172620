// K4 (direct form): non-iterative solver of the per-wavenumber 2-D problems when they are separable
// (SURVEY.md section 8(f) rank 3).
//
// solve_n_helmholtz_2d (reference src/solver.f90:843-922) hands HYPRE one 5-point problem per wavenumber q,
//   [ L_1 + L_2 + lambda_q ] x = b / (w_1 w_2)            (matrix of init_matrix_2d, src/solver.f90:687-842,
//                                                            divided by the cell area w_1 w_2)
// with L_d the 1-D second-difference operator of direction d, boundary conditions folded into its first and
// last row (src/solver.f90:788-823).  When every plane is one rectangle (not split over blocks) and one of its
// two directions, t, has a uniform grid, L_t is diagonalised by the same real-to-real transform the reference
// uses along FFT_AXIS for that boundary-condition pair (find_fft / eigenvalues, src/fft.f90:96-189), and what is
// left is one tridiagonal system along the other direction, s, per (k_t, q):
//   forward transform along t  ->  tridiagonal solves along s  ->  backward transform along t.
// The other direction may be stretched and carry any boundary conditions; when it is periodic the system is
// cyclic and is closed through its last unknown (below).  The factorisation depends only on the grid: it is
// computed once (one array of the size of the slab: 1/pivot; cyclic: a second one with the coupling to the last
// unknown), so a solve moves
//   transforms 2 x 16 B + tridiagonal 48 B (cyclic 56 B) per point,
// against 52 B per point and V-cycle plus the coarse levels for the multigrid (mg.cu), has no stopping test,
// no tail of almost-converged planes and returns the exact solution of the linear system (round-off).
// A singular problem (all-Neumann/periodic plane with lambda_q = 0 and zero transform eigenvalue) is closed by
// x_last = 0, which fixes the free constant; HYPRE leaves it to the initial guess (src/solver.f90:870,908).
#include <algorithm>
#include <cmath>

#include "common.h"

namespace snac {

// per (k, q) column: pivots and, for a cyclic system, the coupling to the last unknown
//   rows j = 1 .. n-1:  x_j + c'_j x_{j+1} + e'_j x_n = d'_j ,  c'_j = ap_j / beta_j
//   beta_j = diag_j - am_j c'_{j-1},  e'_1 = am_1 / beta_1,  e'_j = -am_j e'_{j-1} / beta_j,
//   e'_{n-1} = (ap_{n-1} - am_{n-1} e'_{n-2}) / beta_{n-1}  (and c'_{n-1} = 0)
//   last row: ap_n x_1 + am_n x_{n-1} + diag_n x_n = d_n, reduced with the rows above to g x_n = h
// non-cyclic: the plain Thomas factorisation (e' = 0, g = beta_n).
struct TriArgs {
  int nt, ns, nq;          // transform length, tridiagonal length, planes
  long st_t, st_s, st_q;   // strides of the three directions in the slab arrays
  const double *stab;      // per j (1..ns): {am_j, ap_j, e_j (diagonal of L_s), 1/(w_t w_s[j])}
  const double *mu;        // [nt] eigenvalues of L_t per transform index
  const double *lamq;      // ghosted [nq+2]: lambda_q
  double alpha;
  double *ib, *ie, *ig;    // 1/beta (slab layout), e' (slab layout, cyclic only), 1/g per (k,q)
  int cyclic;
  double out_scale;        // normalisation of the transform pair, applied to the solution
};

__global__ void __launch_bounds__(128) tri_factor_kernel(const TriArgs a) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
  if (k >= a.nt) return;
  const long base = (long)(q + 1) * a.st_q + a.st_s + a.st_t + (long)k * a.st_t;   // element (k, j = 1)
  const double shift = a.alpha + a.lamq[q + 1] + a.mu[k];
  const int n = a.ns;
  double cp = 0.0, ep = 0.0, f = 0.0, g = 0.0, scale = 0.0;
  if (a.cyclic) { f = a.stab[4 * n + 1]; g = a.stab[4 * n + 2] + shift; }   // ap_n, diag_n
  for (int j = 1; j <= (a.cyclic ? n - 1 : n); ++j) {
    const double am = a.stab[4 * j], ap = a.stab[4 * j + 1], dg = a.stab[4 * j + 2] + shift;
    scale = fmax(scale, fabs(dg));
    double beta = dg - am * cp;
    // a singular non-cyclic problem shows up as a vanishing last pivot: x_n = 0 closes it
    double rb = (j == n && fabs(beta) <= 1e-12 * scale) ? 0.0 : 1.0 / beta;
    a.ib[base + (long)(j - 1) * a.st_s] = rb;
    if (a.cyclic) {
      double e;
      if (j == 1) e = am * rb;
      else if (j == n - 1) e = (ap - am * ep) * rb;
      else e = -am * ep * rb;
      a.ie[base + (long)(j - 1) * a.st_s] = e;
      // eliminate x_j from the last row: f is its coefficient there
      g -= f * e;
      const double c = (j == n - 1) ? 0.0 : ap * rb;
      f = -f * c;
      if (j == n - 2) f += a.stab[4 * n];   // + am_n x_{n-1}
      ep = e;
      cp = c;
    } else {
      cp = ap * rb;
    }
  }
  if (a.cyclic) {
    const double dn = fabs(a.stab[4 * n + 2] + shift);
    a.ig[(long)q * a.nt + k] = fabs(g) <= 1e-12 * fmax(dn, scale) ? 0.0 : 1.0 / g;
  }
}

// forward elimination and back substitution of one column per thread; threads of a warp hold consecutive
// transform indices k.  b: transformed right-hand side (unscaled), x: receives d' on the way down and the
// solution on the way up (in place).  The loads of b, 1/beta and e' do not depend on the recurrence: they are
// issued a batch of U rows ahead of their use (explicit register double-buffering), so that every thread keeps
// 2U-3U loads in flight whatever the number of columns -- with few columns (small planes) nothing else hides
// the memory latency.
#ifndef SNAC_TRI_U
#define SNAC_TRI_U 8
#endif
constexpr int TRI_U = SNAC_TRI_U;
template <bool CYC>
__global__ void __launch_bounds__(128) tri_solve_kernel(const TriArgs a, const double *__restrict__ b, double *__restrict__ x) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x, q = blockIdx.y;
  if (k >= a.nt) return;
  const long base = (long)(q + 1) * a.st_q + a.st_s + a.st_t + (long)k * a.st_t;
  const int n = a.ns;
  const long ss = a.st_s;
  const double *__restrict__ ib = a.ib;
  const double *__restrict__ ie = a.ie;
  const double2 *__restrict__ tab2 = reinterpret_cast<const double2 *>(a.stab);   // {am, ap}, {e, 1/(w_t w_s)} per row
  double dp = 0.0, f = 0.0, h = 0.0;
  const int nfw = CYC ? n - 1 : n;
  if (CYC) f = __ldg(&tab2[2 * n]).y;   // ap_n
  double rbv[TRI_U], bv[TRI_U], rbn[TRI_U], bn[TRI_U];
#pragma unroll
  for (int u = 0; u < TRI_U; ++u) {
    const int j = 1 + u;
    const long o = base + (long)(j - 1) * ss;
    rbv[u] = j <= nfw ? ib[o] : 0.0;
    bv[u] = j <= nfw ? b[o] : 0.0;
  }
  for (int j0 = 1; j0 <= nfw; j0 += TRI_U) {
#pragma unroll
    for (int u = 0; u < TRI_U; ++u) {   // the next batch
      const int j = j0 + TRI_U + u;
      const long o = base + (long)(j - 1) * ss;
      rbn[u] = j <= nfw ? ib[o] : 0.0;
      bn[u] = j <= nfw ? b[o] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < TRI_U; ++u) {
      const int j = j0 + u;
      if (j <= nfw) {
        const double2 t = __ldg(&tab2[2 * j]);
        const double rs = __ldg(&tab2[2 * j + 1]).y;
        const double rb = rbv[u];
        const double d = (bv[u] * rs - t.x * dp) * rb;
        x[base + (long)(j - 1) * ss] = d;
        dp = d;
        if (CYC) {
          h -= f * d;
          f = (j == n - 1) ? 0.0 : -f * (t.y * rb);
          if (j == n - 2) f += __ldg(&tab2[2 * n]).x;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < TRI_U; ++u) { rbv[u] = rbn[u]; bv[u] = bn[u]; }
  }
  double xn = 0.0, xp;
  if (CYC) {
    const double rs = __ldg(&tab2[2 * n + 1]).y;
    const long o = base + (long)(n - 1) * ss;
    xn = (b[o] * rs + h) * a.ig[(long)q * a.nt + k];
    x[o] = xn * a.out_scale;
    xp = 0.0;   // c'_{n-1} = 0: row n-1 couples to x_n through e' only
  } else {
    xp = dp;    // x_n = d'_n
    x[base + (long)(n - 1) * ss] = xp * a.out_scale;
  }
  // back substitution, rows jtop .. 1; x holds d' there (written by this thread above)
  const int jtop = nfw - (CYC ? 0 : 1);
  double ev[TRI_U], en[TRI_U];
#pragma unroll
  for (int u = 0; u < TRI_U; ++u) {
    const int j = jtop - u;
    const long o = base + (long)(j - 1) * ss;
    rbv[u] = j >= 1 ? ib[o] : 0.0;
    bv[u] = j >= 1 ? x[o] : 0.0;
    if (CYC) ev[u] = j >= 1 ? ie[o] : 0.0;
  }
  for (int j0 = jtop; j0 >= 1; j0 -= TRI_U) {
#pragma unroll
    for (int u = 0; u < TRI_U; ++u) {
      const int j = j0 - TRI_U - u;
      const long o = base + (long)(j - 1) * ss;
      rbn[u] = j >= 1 ? ib[o] : 0.0;
      bn[u] = j >= 1 ? x[o] : 0.0;
      if (CYC) en[u] = j >= 1 ? ie[o] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < TRI_U; ++u) {
      const int j = j0 - u;
      if (j >= 1) {
        const double cpj = (CYC && j == n - 1) ? 0.0 : __ldg(&tab2[2 * j]).y * rbv[u];
        double v = bv[u] - cpj * xp;
        if (CYC) v -= ev[u] * xn;
        x[base + (long)(j - 1) * ss] = v * a.out_scale;
        xp = v;
      }
    }
#pragma unroll
    for (int u = 0; u < TRI_U; ++u) { rbv[u] = rbn[u]; bv[u] = bn[u]; if (CYC) ev[u] = en[u]; }
  }
}

static bool uniform(const std::vector<double> &v, int n) {   // entries 1..n equal (to round-off)
  for (int i = 2; i <= n; ++i)
    if (std::fabs(v[i] - v[1]) > 1e-12 * std::fabs(v[1])) return false;
  return true;
}

// Can direction t of solver box sb be transformed?  Uniform widths and spacings (ghost entries included where a
// neighbour or a wall image is coupled), and a boundary-condition pair with a transform.
static bool transformable(const SBox &sb, int t, int self, char bc[2]) {
  const int n = sb.n[t];
  if (n < 2 || !uniform(sb.d2[t], n)) return false;
  const bool per = sb.nb[t][0] >= 0 || sb.nb[t][1] >= 0;
  if (per) {
    if (sb.nb[t][0] != self || sb.nb[t][1] != self) return false;
    for (int i = 0; i <= n; ++i)
      if (std::fabs(sb.d1[t][i] - sb.d2[t][1]) > 1e-12 * sb.d2[t][1]) return false;
    bc[0] = bc[1] = 'F';
    return true;
  }
  if (!sb.phys[t][0] || !sb.phys[t][1]) return false;
  for (int i = 1; i < n; ++i)
    if (std::fabs(sb.d1[t][i] - sb.d2[t][1]) > 1e-12 * sb.d2[t][1]) return false;
  for (int ib = 0; ib < 2; ++ib) bc[ib] = sb.sgn[t][ib] > 0.0 ? 'N' : (sb.sgn[t][ib] < 0.0 ? 'D' : '?');
  return bc[0] != '?' && bc[1] != '?';
}

// Which of the two plane directions is transformed (-1: the planes are not separable).  Prefers direction 0
// (unit stride: the tridiagonal kernel is coalesced over the transform index) when both qualify.
int direct_pick(const std::vector<SBox> &sboxes, bool ids_local, const std::vector<int> &local_of_rank) {
  int best = -1;
  for (int t = 1; t >= 0; --t) {
    bool ok = true;
    for (size_t lb = 0; lb < sboxes.size() && ok; ++lb) {
      const SBox &sb = sboxes[lb];
      const int s = 1 - t;
      char bc[2];
      if (!transformable(sb, t, (int)lb, bc)) ok = false;
      // the other direction: closed by physical faces or periodic onto itself (not split over boxes)
      for (int ib = 0; ib < 2 && ok; ++ib) {
        const int nb = sb.nb[s][ib];
        if (nb >= 0 && (ids_local ? nb : local_of_rank[nb]) != (int)lb) ok = false;
        if (nb < 0 && !sb.phys[s][ib]) ok = false;
      }
      if ((sb.nb[s][0] >= 0) != (sb.nb[s][1] >= 0)) ok = false;
      if (sb.nb[s][0] >= 0 && sb.n[s] < 3) ok = false;
      if (sb.qqq[0] || sb.qqq[1]) ok = false;   // centred unknowns only
    }
    if (ok) best = t;
  }
  return best;
}

void DirectSolver::setup(Ctx &ctx, MGSolver &S, int tdir_) {
  tdir = tdir_;
  const int sdir = 1 - tdir;
  const int nl = (int)S.sboxes.size();
  box.clear();
  box.resize(nl);
  for (int lb = 0; lb < nl; ++lb) {
    const SBox &sb = S.sboxes[lb];
    DirectBox &B = box[lb];
    const Grid3 &g = S.grid0(lb);
    char bc[2];
    if (!transformable(sb, tdir, lb, bc)) throw Error("direct solver: direction is not transformable");
    r2r_plan_init(B.plan, tdir, sb.n[tdir], bc, true);
    const double h = sb.d2[tdir][1];
    std::vector<double> mu = r2r_eigenvalues(sb.n[tdir], bc, true);
    for (double &m : mu) m /= h * h;
    B.mu.upload(mu);
    // 1-D operator of the other direction exactly as the multigrid builds it (build_dir, mg.cu): am = 1/d1[j-1],
    // ap = 1/d1[j], ad = -(am + ap), physical faces folded into the diagonal; scaled by 1/w
    const int n = sb.n[sdir];
    B.cyclic = sb.nb[sdir][0] >= 0;
    std::vector<double> tab(4 * (size_t)(n + 2), 0.0);
    for (int j = 1; j <= n; ++j) {
      double am = 1.0 / sb.d1[sdir][j - 1], ap = 1.0 / sb.d1[sdir][j], ad = -(am + ap);
      if (j == 1 && sb.phys[sdir][0]) { ad += sb.sgn[sdir][0] * am; am = 0.0; }
      if (j == n && sb.phys[sdir][1]) { ad += sb.sgn[sdir][1] * ap; ap = 0.0; }
      const double rw = 1.0 / sb.d2[sdir][j];
      tab[4 * j] = am * rw; tab[4 * j + 1] = ap * rw; tab[4 * j + 2] = ad * rw; tab[4 * j + 3] = rw / h;
    }
    B.stab.upload(tab);
    B.ib.alloc(g.alloc_elems());
    if (B.cyclic) { B.ie.alloc(g.alloc_elems()); B.ig.alloc((size_t)sb.n[tdir] * S.nq); }
    TriArgs a = args(S, lb, 0.0);
    dim3 grd((a.nt + 127) / 128, S.nq);
    tri_factor_kernel<<<grd, 128, 0, ctx.stream>>>(a);
    SNAC_CUDA(cudaGetLastError());
  }
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
  on = true;
}

TriArgs DirectSolver::args(const MGSolver &S, int lb, double alpha) const {
  const DirectBox &B = box[lb];
  const Grid3 &g = S.grid0(lb);
  TriArgs a;
  a.nt = g.n[tdir]; a.ns = g.n[1 - tdir]; a.nq = S.nq;
  a.st_t = g.st(tdir); a.st_s = g.st(1 - tdir); a.st_q = g.sz;
  a.stab = B.stab.p(); a.mu = B.mu.p(); a.lamq = S.lam[lb].p();
  a.alpha = alpha;
  a.ib = B.ib.raw + 15; a.ie = B.cyclic ? B.ie.raw + 15 : nullptr; a.ig = B.ig.raw;
  a.cyclic = B.cyclic ? 1 : 0;
  a.out_scale = B.plan.normfft;
  return a;
}

void DirectSolver::solve(Ctx &ctx, MGSolver &S) {
  const int nl = (int)S.sboxes.size();
  for (int lb = 0; lb < nl; ++lb) {
    DirectBox &B = box[lb];
    const Grid3 &g = S.grid0(lb);
    double *x = S.xcur(lb), *b = S.bcur(lb);
    const int n3[3] = {g.n[0], g.n[1], S.nq};
    const long st[3] = {1, g.sy, g.sz};
    const long pint = 1 + g.sy + g.sz;
    const double pts = (double)g.n[0] * g.n[1] * S.nq;
    {
      ProfScope ps(ctx, PC_DIRECT_TRANSFORM, pts);
      const double *in[1] = {b + pint};
      double *out[1] = {b + pint};
      r2r_execute(ctx, B.plan, +1, n3, in, 1, st, out, 1, st, 1, 1.0);
    }
    {
      ProfScope ps(ctx, PC_DIRECT_TRIDIAG, pts);
      TriArgs a = args(S, lb, S.alpha);
      dim3 grd((a.nt + 127) / 128, S.nq);
      if (B.cyclic) tri_solve_kernel<true><<<grd, 128, 0, ctx.stream>>>(a, b, x);
      else tri_solve_kernel<false><<<grd, 128, 0, ctx.stream>>>(a, b, x);
      ctx.launches++;
      SNAC_CUDA(cudaGetLastError());
    }
    {
      ProfScope ps(ctx, PC_DIRECT_TRANSFORM, pts);
      const double *in[1] = {x + pint};
      double *out[1] = {x + pint};
      r2r_execute(ctx, B.plan, -1, n3, in, 1, st, out, 1, st, 1, 1.0);
    }
  }
  S.last_iters = 1;
  S.last_relres = 0.0;
  S.last_mean_iters = 1.0;
}

}  // namespace snac
