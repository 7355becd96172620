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
// plane-local fast path (batched 2-D problems whose planes are not split over boxes)
//
// Canonical layout: directions 0,1 are the coupled ones (0 fastest), direction 2 is the batch
// (wavenumber).  A V(1,1) cycle touches each fine level with two fused kernels,
//   down: red sweep + black sweep + residual + restriction      (read x, b; write x', coarse b)
//   up  : correction from the coarse level + black sweep + red sweep   (read x', b, coarse x; write x)
// working on shared-memory tiles with a halo wide enough to recompute the neighbours' updates
// (3 cells down, 2 cells up); x is ping-ponged between two buffers so that tiles never read what
// another tile has already overwritten.  All levels whose plane fits in shared memory are done by
// one kernel (one CTA per wavenumber) that runs the rest of the V-cycle without leaving the SM.
// ------------------------------------------------------------------------------------------
struct PLevel {
  int n1, n2;
  long sy, sz;        // row / plane strides of this level's global arrays
  const double *am1, *ap1, *ad1, *w1, *am2, *ap2, *ad2, *w2;  // ghosted, index 1..n
  const double *sam1, *sap1, *se1, *srw1, *sam2, *sap2, *se2, *srw2;  // am/w, ap/w, ad/w, 1/w
  int c1, c2;         // coarsening towards the next level
  int per1, per2;     // periodic wrap onto the box itself
};

// Tile geometry: 64 columns including the halo, stored split by colour so that each lane owns one
// column pair (2k, 2k+1) and every shared-memory access of a sweep is stride-1 across the warp;
// a warp works on two rows at a time (so that the column parity of a colour is a compile-time
// constant), the 8 warps of a CTA on row pairs 16 apart.
constexpr int PW = 64, PHW = PW / 2;
constexpr int PT_Y = 32;                // interior rows of a tile
constexpr int PH_D = 3, PH_U = 2;       // halo of the down / up kernels
__host__ __device__ constexpr int ptx(int H) { return PW - 2 * H; }   // interior columns of a tile

__device__ inline int wrapn(int g, int n) {  // 1-based periodic wrap
  int r = (g - 1) % n;
  if (r < 0) r += n;
  return r + 1;
}

// The cell equations are used in the scaled form (division by w1*w2):
//   x = ( b/(w1 w2) - (a1m xL + a1p xR + a2m xD + a2p xU) ) / (e1 + e2 + lambda)
// with a = am/w, e = ad/w precomputed per level on the host (PLevel::s*).  Cells outside a
// non-periodic domain get a = 0 and e = +inf, so that "updating" them yields exactly 0 without a
// branch; halo cells too close to the tile edge are updated with incomplete data, which is
// harmless because nothing that is kept reads them.
template <int H>
struct PTileSmem {
  static constexpr int HT = PT_Y + 2 * H;
  double x[2][HT][PHW];
  double b[2][HT][PHW];
  double ram[HT], rap[HT], re[HT], rw[HT], rrw[HT];
  int rgj[HT];   // global row index, 0: outside the domain
};

struct PCols {   // the two columns (ti = 2*lane, 2*lane+1) owned by a lane; e includes lambda
  double am0, am1, ap0, ap1, e0, e1, w0, w1, rw0, rw1;
  int gi0, gi1;  // global column index, 0: outside the domain
};

template <int H>
__device__ inline void pt_setup(const PLevel &L, int i0, int j0, double lam, PCols &C, PTileSmem<H> &S) {
  constexpr int HT = PT_Y + 2 * H;
  const int lane = threadIdx.x & 31;
  auto col = [&](int sft, double &am, double &ap, double &e, double &w, double &rw, int &gi) {
    const int gu = i0 + 2 * lane + sft - H;
    const bool ok = L.per1 || (gu >= 1 && gu <= L.n1);
    gi = ok ? (L.per1 ? wrapn(gu, L.n1) : gu) : 0;
    am = ok ? __ldg(L.sam1 + gi) : 0.0; ap = ok ? __ldg(L.sap1 + gi) : 0.0;
    e = ok ? __ldg(L.se1 + gi) + lam : CUDART_INF;
    w = ok ? __ldg(L.w1 + gi) : 0.0; rw = ok ? __ldg(L.srw1 + gi) : 0.0;
  };
  col(0, C.am0, C.ap0, C.e0, C.w0, C.rw0, C.gi0);
  col(1, C.am1, C.ap1, C.e1, C.w1, C.rw1, C.gi1);
  for (int t = threadIdx.x; t < HT; t += blockDim.x) {
    const int gv = j0 + t - H;
    const bool ok = L.per2 || (gv >= 1 && gv <= L.n2);
    const int gj = ok ? (L.per2 ? wrapn(gv, L.n2) : gv) : 0;
    S.rgj[t] = gj;
    S.ram[t] = ok ? __ldg(L.sam2 + gj) : 0.0; S.rap[t] = ok ? __ldg(L.sap2 + gj) : 0.0;
    S.re[t] = ok ? __ldg(L.se2 + gj) : CUDART_INF;
    S.rw[t] = ok ? __ldg(L.w2 + gj) : 0.0; S.rrw[t] = ok ? __ldg(L.srw2 + gj) : 0.0;
  }
}

// update of the cell (row tj, column 2*lane + SFT) of colour COLOR
template <int H, int COLOR, int SFT>
__device__ __forceinline__ void pt_update(PTileSmem<H> &S, const PCols &C, int tj, int lane) {
  const double *xo = &S.x[1 - COLOR][0][0];
  const int c = tj * PHW + lane;
  const double am = SFT ? C.am1 : C.am0, ap = SFT ? C.ap1 : C.ap0, e = SFT ? C.e1 : C.e0;
  const double off = am * xo[c - 1 + SFT] + ap * xo[c + SFT] + S.ram[tj] * xo[c - PHW] + S.rap[tj] * xo[c + PHW];
  S.x[COLOR][tj][lane] = (S.b[COLOR][tj][lane] - off) * __drcp_rn(e + S.re[tj]);
}

// one colour of a Gauss-Seidel sweep on rows RING .. HT-RING-1 (all columns, see above)
template <int H, int RING, int COLOR>
__device__ __forceinline__ void pt_sweep(PTileSmem<H> &S, const PCols &C) {
  constexpr int HT = PT_Y + 2 * H;
  constexpr int S0 = (COLOR + RING) & 1;   // column parity of this colour in row RING
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int tj = RING + 2 * warp; tj < HT - RING; tj += 16) {
    pt_update<H, COLOR, S0>(S, C, tj, lane);
    if (tj + 1 < HT - RING) pt_update<H, COLOR, 1 - S0>(S, C, tj + 1, lane);
  }
}

__global__ void __launch_bounds__(256) mg_plane_down_kernel(const PLevel L, const double *xin, double *xout, const double *b,
                                                            double *cb, long csy, long csz, const double *lamq, double alpha,
                                                            double *norms, int nq, int want_b, const int *active) {
  constexpr int H = PH_D, HT = PT_Y + 2 * H, TX = ptx(H);
  static_assert(H == 3, "the fused restriction assumes an odd halo width");
  __shared__ PTileSmem<H> S;
  __shared__ double red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.z;
  if (active && !active[q]) return;   // this wavenumber has converged
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * PT_Y;
  const long qoff = (long)(q + 1) * L.sz;
  const double lam = alpha + __ldg(lamq + q + 1);
  PCols C;
  pt_setup<H>(L, i0, j0, lam, C, S);
  __syncthreads();
  double r2 = 0.0, b2 = 0.0;
  // load x and the scaled right-hand side, split by colour
  for (int tj = warp; tj < HT; tj += 8) {
    const int gj = S.rgj[tj];
    const double rrw = S.rrw[tj];
    const bool rin = tj >= H && tj < H + PT_Y && (j0 + tj - H) <= L.n2;
    const long orow = qoff + L.sy * gj;
    double v0 = 0.0, v1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (gj > 0) {
      if (C.gi0 > 0) { if (xin) v0 = xin[orow + C.gi0]; b0 = b[orow + C.gi0]; }
      if (C.gi1 > 0) { if (xin) v1 = xin[orow + C.gi1]; b1 = b[orow + C.gi1]; }
    }
    if (rin) {
      if (2 * lane >= H && 2 * lane < PW - H && (i0 + 2 * lane - H) <= L.n1) b2 += b0 * b0;
      if (2 * lane + 1 >= H && 2 * lane + 1 < PW - H && (i0 + 2 * lane + 1 - H) <= L.n1) b2 += b1 * b1;
    }
    const int c0 = tj & 1;   // colour of column 2*lane in this row
    S.x[c0][tj][lane] = v0; S.b[c0][tj][lane] = b0 * (C.rw0 * rrw);
    S.x[1 - c0][tj][lane] = v1; S.b[1 - c0][tj][lane] = b1 * (C.rw1 * rrw);
  }
  __syncthreads();
  pt_sweep<H, 1, 0>(S, C);
  __syncthreads();
  pt_sweep<H, 2, 1>(S, C);
  __syncthreads();
  // new iterate to global memory, residual (non-zero on colour 0 only) and its restriction
  const bool full = L.c1 && L.c2;
  const int cn1 = L.c1 ? (L.n1 + 1) / 2 : L.n1, cn2 = L.c2 ? (L.n2 + 1) / 2 : L.n2;
  for (int tj = H + 2 * warp; tj < H + PT_Y; tj += 16) {   // tj is odd: (tj, tj+1) are the children rows of a coarse row
    double resA = 0.0, resB = 0.0;   // residual at (tj, column 2*lane+1) and at (tj+1, column 2*lane)
#pragma unroll
    for (int dr = 0; dr < 2; ++dr) {
      const int t = tj + dr, gv = j0 + t - H;
      const bool rin = gv <= L.n2;
      const int gu0 = i0 + 2 * lane - H, gu1 = gu0 + 1;
      const bool in0 = rin && 2 * lane >= H && 2 * lane < PW - H && gu0 <= L.n1;
      const bool in1 = rin && 2 * lane + 1 >= H && 2 * lane + 1 < PW - H && gu1 <= L.n1;
      const int c0 = t & 1;
      const long orow = qoff + L.sy * gv;
      if (in0) xout[orow + gu0] = S.x[c0][t][lane];
      if (in1) xout[orow + gu1] = S.x[1 - c0][t][lane];
      // colour-0 cell of this row: column parity = row parity (dr = 0: odd row -> column 2*lane+1)
      const double *xo = &S.x[1][0][0];
      const int c = t * PHW + lane;
      if (dr == 0) {
        if (in1) {
          const double off = C.am1 * xo[c] + C.ap1 * xo[c + 1] + S.ram[t] * xo[c - PHW] + S.rap[t] * xo[c + PHW];
          resA = (S.b[0][t][lane] - off - (C.e1 + S.re[t]) * S.x[0][t][lane]) * (C.w1 * S.rw[t]);
        }
      } else {
        if (in0) {
          const double off = C.am0 * xo[c - 1] + C.ap0 * xo[c] + S.ram[t] * xo[c - PHW] + S.rap[t] * xo[c + PHW];
          resB = (S.b[0][t][lane] - off - (C.e0 + S.re[t]) * S.x[0][t][lane]) * (C.w0 * S.rw[t]);
        }
      }
    }
    r2 += resA * resA + resB * resB;
    if (full) {
      // coarse cell = columns (2*lane+1, 2*lane+2) x rows (tj, tj+1): its colour-0 children are
      // (tj, 2*lane+1) on this lane and (tj+1, 2*lane+2) on the next one
      const double nxt = __shfl_down_sync(0xffffffffu, resB, 1);
      const int I = (i0 - 1) / 2 + lane, J = (j0 + tj - H + 1) / 2;
      if (cb && lane >= 1 && lane <= (PW - H - 2) / 2 && I <= cn1 && J <= cn2)
        cb[(long)(q + 1) * csz + I + csy * J] = resA + nxt;
    } else {
      S.b[0][tj][lane] = resA;       // (tj odd: colour 0 sits in column 2*lane+1; tj+1: column 2*lane)
      S.b[0][tj + 1][lane] = resB;
    }
  }
  if (!full && cb) {   // semi-coarsening: generic restriction through shared memory
    __syncthreads();
    const int cw = L.c1 ? TX / 2 : TX, ch = L.c2 ? PT_Y / 2 : PT_Y;
    const int I0 = L.c1 ? (i0 + 1) / 2 : i0, J0 = L.c2 ? (j0 + 1) / 2 : j0;
    for (int t = threadIdx.x; t < cw * ch; t += blockDim.x) {
      const int ci = t % cw, cj = t / cw;
      const int I = I0 + ci, J = J0 + cj;
      if (I > cn1 || J > cn2) continue;
      const int ti0 = H + (L.c1 ? 2 * ci : ci), tj0 = H + (L.c2 ? 2 * cj : cj);
      double sum = 0.0;
      for (int dj = 0; dj <= L.c2; ++dj)
        for (int di = 0; di <= L.c1; ++di) {
          const int ti = ti0 + di, tj = tj0 + dj;
          if (((ti + tj) & 1) == 0) sum += S.b[0][tj][ti >> 1];
        }
      cb[(long)(q + 1) * csz + I + csy * J] = sum;
    }
  }
  if (norms) {
    for (int o = 16; o > 0; o >>= 1) {
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    if (lane == 0) { red[0][warp] = r2; red[1][warp] = b2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) { r2 += red[0][w]; b2 += red[1][w]; }
      atomicAdd(norms + q, r2);
      if (want_b) atomicAdd(norms + nq + q, b2);
    }
  }
}

__global__ void __launch_bounds__(256) mg_plane_up_kernel(const PLevel L, const double *xin, double *xout, const double *b,
                                                          const double *cx, long csy, long csz, const double *lamq,
                                                          double alpha, const int *active) {
  constexpr int H = PH_U, HT = PT_Y + 2 * H, TX = ptx(H);
  __shared__ PTileSmem<H> S;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.z;
  if (active && !active[q]) return;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * PT_Y;
  const long qoff = (long)(q + 1) * L.sz, cqoff = (long)(q + 1) * csz;
  const double lam = alpha + __ldg(lamq + q + 1);
  PCols C;
  pt_setup<H>(L, i0, j0, lam, C, S);
  __syncthreads();
  // load x + coarse-grid correction (piecewise-constant prolongation) and the scaled right-hand side
  const int I0 = L.c1 ? (C.gi0 + 1) >> 1 : C.gi0, I1 = L.c1 ? (C.gi1 + 1) >> 1 : C.gi1;
  for (int tj = warp; tj < HT; tj += 8) {
    const int gj = S.rgj[tj];
    const double rrw = S.rrw[tj];
    const long orow = qoff + L.sy * gj;
    const long crow = cqoff + csy * (L.c2 ? (gj + 1) >> 1 : gj);
    double v0 = 0.0, v1 = 0.0, b0 = 0.0, b1 = 0.0;
    if (gj > 0) {
      if (C.gi0 > 0) { v0 = xin[orow + C.gi0] + __ldg(cx + crow + I0); b0 = b[orow + C.gi0]; }
      if (C.gi1 > 0) { v1 = xin[orow + C.gi1] + __ldg(cx + crow + I1); b1 = b[orow + C.gi1]; }
    }
    const int c0 = tj & 1;
    S.x[c0][tj][lane] = v0; S.b[c0][tj][lane] = b0 * (C.rw0 * rrw);
    S.x[1 - c0][tj][lane] = v1; S.b[1 - c0][tj][lane] = b1 * (C.rw1 * rrw);
  }
  __syncthreads();
  pt_sweep<H, 1, 1>(S, C);
  __syncthreads();
  pt_sweep<H, 2, 0>(S, C);
  __syncthreads();
  for (int tj = H + warp; tj < H + PT_Y; tj += 8) {
    const int gv = j0 + tj - H;
    if (gv > L.n2) continue;
    const int gu0 = i0 + 2 * lane - H, gu1 = gu0 + 1;
    const int c0 = tj & 1;
    const long orow = qoff + L.sy * gv;
    if (2 * lane >= H && 2 * lane < PW - H && gu0 <= L.n1) xout[orow + gu0] = S.x[c0][tj][lane];
    if (2 * lane + 1 >= H && 2 * lane + 1 < PW - H && gu1 <= L.n1) xout[orow + gu1] = S.x[1 - c0][tj][lane];
  }
}

// ------------------------------------------------------------------------------------------
// Register-marching variants of the two fused kernels.  A warp owns 64 columns (two per lane) and
// marches over the rows of its chunk, keeping a window of six rows in registers; the horizontal
// neighbours of a cell are the lane's other column or one warp shuffle away, the vertical ones are
// other window slots, so the stencil needs no shared memory and no block-level barrier.  Row t is
// loaded one step ahead; at step t the first colour of row t-1, the second colour of row t-2 and
// the output (store, residual, restriction) of row t-3 (down) / t-2 (up) are done.
// ------------------------------------------------------------------------------------------
constexpr int MRY = 32;   // interior rows marched by one warp

struct MCols {            // per-lane column data (columns 2*lane, 2*lane+1 of the warp's strip)
  double am0, am1, ap0, ap1, e0, e1, w0, w1, rw0, rw1;
  int gi0, gi1;
};

template <int H>
__device__ __forceinline__ void march_cols(const PLevel &L, int i0, double lam, int lane, MCols &C) {
  auto col = [&](int sft, double &am, double &ap, double &e, double &w, double &rw, int &gi) {
    const int gu = i0 + 2 * lane + sft - H;
    const bool ok = L.per1 || (gu >= 1 && gu <= L.n1);
    gi = ok ? (L.per1 ? wrapn(gu, L.n1) : gu) : 0;
    am = ok ? __ldg(L.sam1 + gi) : 0.0; ap = ok ? __ldg(L.sap1 + gi) : 0.0;
    e = ok ? __ldg(L.se1 + gi) + lam : CUDART_INF;
    w = ok ? __ldg(L.w1 + gi) : 0.0; rw = ok ? __ldg(L.srw1 + gi) : 0.0;
  };
  col(0, C.am0, C.ap0, C.e0, C.w0, C.rw0, C.gi0);
  col(1, C.am1, C.ap1, C.e1, C.w1, C.rw1, C.gi1);
}

// per-warp row table in shared memory: [0] a2m, [1] a2p, [2] e2, [3] w2, [4] 1/w2, [5] global row (as double; 0: outside)
template <int H>
__device__ __forceinline__ void march_rows(const PLevel &L, int j0, int lane, double (*rt)[MRY + 2 * H]) {
  constexpr int HT = MRY + 2 * H;
  for (int t = lane; t < HT; t += 32) {
    const int gv = j0 + t - H;
    const bool ok = L.per2 || (gv >= 1 && gv <= L.n2);
    const int gj = ok ? (L.per2 ? wrapn(gv, L.n2) : gv) : 0;
    rt[0][t] = ok ? __ldg(L.sam2 + gj) : 0.0; rt[1][t] = ok ? __ldg(L.sap2 + gj) : 0.0;
    rt[2][t] = ok ? __ldg(L.se2 + gj) : CUDART_INF;
    rt[3][t] = ok ? __ldg(L.w2 + gj) : 0.0; rt[4][t] = ok ? __ldg(L.srw2 + gj) : 0.0;
    rt[5][t] = (double)gj;
  }
  __syncwarp();
}

// Gauss-Seidel update of one cell.  A: the cell is in column 2*lane (else 2*lane+1); xs/xo: this
// row's own-column and other-column values; xm/xp: the own column one row below/above.
#define MARCH_UPD(A, xs, xo, xm, xp, bs, r)                                                        \
  {                                                                                                 \
    const double l_ = (A) ? __shfl_up_sync(0xffffffffu, xo, 1) : xo;                                \
    const double r_ = (A) ? xo : __shfl_down_sync(0xffffffffu, xo, 1);                              \
    const double off_ = ((A) ? C.am0 : C.am1) * l_ + ((A) ? C.ap0 : C.ap1) * r_ + rt[0][r] * (xm) + rt[1][r] * (xp); \
    xs = ((bs) - off_) * __drcp_rn(((A) ? C.e0 : C.e1) + rt[2][r]);                                 \
  }

__global__ void __launch_bounds__(256) mg_march_down_kernel(const PLevel L, const double *__restrict__ xin, double *__restrict__ xout,
                                                            const double *__restrict__ b, double *__restrict__ cb, long csy, long csz,
                                                            const double *lamq, double alpha, double *norms, int nq, int want_b,
                                                            int ntx, int nty, const int *active) {
  constexpr int H = 3, HT = MRY + 2 * H, TX = PW - 2 * H;
  __shared__ double rtab[8][6][HT];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long wid = (long)blockIdx.x * 8 + wib;
  const int tx = (int)(wid % ntx), ty = (int)((wid / ntx) % nty), q = (int)(wid / ((long)ntx * nty));
  if (q >= nq || (active && !active[q])) return;
  double (*rt)[HT] = rtab[wib];
  const int i0 = 1 + tx * TX, j0 = 1 + ty * MRY;
  const long qoff = (long)(q + 1) * L.sz;
  const double lam = alpha + __ldg(lamq + q + 1);
  MCols C;
  march_cols<H>(L, i0, lam, lane, C);
  march_rows<H>(L, j0, lane, rt);
  const bool cin0 = 2 * lane >= H && 2 * lane < PW - H && (i0 + 2 * lane - H) <= L.n1;          // column kept by this warp
  const bool cin1 = 2 * lane + 1 >= H && 2 * lane + 1 < PW - H && (i0 + 2 * lane + 1 - H) <= L.n1;
  const int gu0 = i0 + 2 * lane - H;
  const int cn1 = L.c1 ? (L.n1 + 1) / 2 : L.n1, cn2 = L.c2 ? (L.n2 + 1) / 2 : L.n2;
  const int Ic = (i0 - 1) / 2 + lane;   // coarse column of the pair (2*lane+1, 2*lane+2) when c1
  const bool cpair = lane >= 1 && lane <= (PW - H - 2) / 2 && Ic <= cn1;
  const long cq = (long)(q + 1) * csz;
  double xa[6], xb[6], ba[6], bb[6];
  double r2 = 0.0, b2 = 0.0, resA = 0.0;

#define MARCH_LOAD(t, S)                                                                           \
  {                                                                                                 \
    const int gj_ = (int)rt[5][t];                                                                  \
    double v0_ = 0.0, v1_ = 0.0, b0_ = 0.0, b1_ = 0.0;                                              \
    if (gj_ > 0) {                                                                                  \
      const long o_ = qoff + L.sy * gj_;                                                            \
      if (C.gi0 > 0) { if (xin) v0_ = xin[o_ + C.gi0]; b0_ = b[o_ + C.gi0]; }                       \
      if (C.gi1 > 0) { if (xin) v1_ = xin[o_ + C.gi1]; b1_ = b[o_ + C.gi1]; }                       \
    }                                                                                               \
    if (want_b && (t) >= H && (t) < H + MRY && (j0 + (t) - H) <= L.n2)                              \
      b2 += (cin0 ? b0_ * b0_ : 0.0) + (cin1 ? b1_ * b1_ : 0.0);                                    \
    const double s_ = rt[4][t];                                                                     \
    xa[S] = v0_; xb[S] = v1_; ba[S] = b0_ * (C.rw0 * s_); bb[S] = b1_ * (C.rw1 * s_);               \
  }

  // one step: K = t mod 6 (so that window slots and the row parity are compile-time constants)
#define MARCH_DOWN_STEP(K)                                                                          \
  if (t0 + (K) < HT + 0) {                                                                          \
    const int t = t0 + (K);                                                                         \
    if (t + 1 < HT) MARCH_LOAD(t + 1, ((K) + 1) % 6)                                                \
    if (t >= 2) { /* colour 0 of row t-1 (slot K+5): column 2*lane in even rows */                  \
      constexpr int R = ((K) + 5) % 6, Rm = ((K) + 4) % 6, Rp = (K);                                \
      if ((((K) + 1) & 1) == 0) MARCH_UPD(1, xa[R], xb[R], xa[Rm], xa[Rp], ba[R], t - 1)            \
      else MARCH_UPD(0, xb[R], xa[R], xb[Rm], xb[Rp], bb[R], t - 1)                                 \
    }                                                                                               \
    if (t >= 4) { /* colour 1 of row t-2 (slot K+4): column 2*lane+1 in even rows */                \
      constexpr int R = ((K) + 4) % 6, Rm = ((K) + 3) % 6, Rp = ((K) + 5) % 6;                      \
      if (((K) & 1) == 0) MARCH_UPD(0, xb[R], xa[R], xb[Rm], xb[Rp], bb[R], t - 2)                  \
      else MARCH_UPD(1, xa[R], xb[R], xa[Rm], xa[Rp], ba[R], t - 2)                                 \
    }                                                                                               \
    if (t >= 2 * H) { /* output of row t-3 (slot K+3) */                                            \
      constexpr int R = ((K) + 3) % 6, Rm = ((K) + 2) % 6, Rp = ((K) + 4) % 6;                      \
      const int r = t - 3, gv = j0 + r - H;                                                         \
      const bool rin = gv <= L.n2;                                                                  \
      const long o = qoff + L.sy * gv;                                                              \
      if (rin && cin0) xout[o + gu0] = xa[R];                                                       \
      if (rin && cin1) xout[o + gu0 + 1] = xb[R];                                                   \
      if ((((K) + 1) & 1) == 1) { /* odd row: colour 0 in column 2*lane+1 */                        \
        const double rgt = __shfl_down_sync(0xffffffffu, xa[R], 1);                                 \
        double res = (bb[R] - (C.am1 * xa[R] + C.ap1 * rgt + rt[0][r] * xb[Rm] + rt[1][r] * xb[Rp]) - \
                      (C.e1 + rt[2][r]) * xb[R]) * (C.w1 * rt[3][r]);                               \
        if (!(rin && cin1)) res = 0.0;                                                              \
        r2 += res * res;                                                                            \
        resA = res;                                                                                 \
        if (cb && L.c1 && !L.c2 && cpair && gv <= cn2) cb[cq + Ic + csy * gv] = res;                \
      } else { /* even row: colour 0 in column 2*lane */                                            \
        const double lft = __shfl_up_sync(0xffffffffu, xb[R], 1);                                   \
        double res = (ba[R] - (C.am0 * lft + C.ap0 * xb[R] + rt[0][r] * xa[Rm] + rt[1][r] * xa[Rp]) - \
                      (C.e0 + rt[2][r]) * xa[R]) * (C.w0 * rt[3][r]);                               \
        if (!(rin && cin0)) res = 0.0;                                                              \
        r2 += res * res;                                                                            \
        if (cb) {                                                                                   \
          const int J = (gv + 1) >> 1;   /* (gv-1, gv) are the children rows of coarse row J */     \
          if (L.c1) {                                                                               \
            const double nxt = __shfl_down_sync(0xffffffffu, res, 1);                               \
            if (L.c2) { if (cpair && J <= cn2) cb[cq + Ic + csy * J] = resA + nxt; }                \
            else if (cpair && gv <= cn2) cb[cq + Ic + csy * gv] = nxt;                              \
          } else if (L.c2 && J <= cn2) {                                                            \
            if (cin0) cb[cq + gu0 + csy * J] = res;                                                 \
            if (cin1) cb[cq + gu0 + 1 + csy * J] = resA;                                            \
          }                                                                                         \
        }                                                                                           \
      }                                                                                             \
    }                                                                                               \
  }

  MARCH_LOAD(0, 0)
  for (int t0 = 0; t0 < HT; t0 += 6) {
    MARCH_DOWN_STEP(0) MARCH_DOWN_STEP(1) MARCH_DOWN_STEP(2) MARCH_DOWN_STEP(3) MARCH_DOWN_STEP(4) MARCH_DOWN_STEP(5)
  }
#undef MARCH_DOWN_STEP
#undef MARCH_LOAD
  if (norms) {
    for (int o = 16; o > 0; o >>= 1) {
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    if (lane == 0) {
      atomicAdd(norms + q, r2);
      if (want_b) atomicAdd(norms + nq + q, b2);
    }
  }
}

__global__ void __launch_bounds__(256) mg_march_up_kernel(const PLevel L, const double *__restrict__ xin, double *__restrict__ xout,
                                                          const double *__restrict__ b, const double *__restrict__ cx, long csy,
                                                          long csz, const double *lamq, double alpha, int nq, int ntx, int nty,
                                                          const int *active) {
  constexpr int H = 2, HT = MRY + 2 * H, TX = PW - 2 * H;
  __shared__ double rtab[8][6][HT];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long wid = (long)blockIdx.x * 8 + wib;
  const int tx = (int)(wid % ntx), ty = (int)((wid / ntx) % nty), q = (int)(wid / ((long)ntx * nty));
  if (q >= nq || (active && !active[q])) return;
  double (*rt)[HT] = rtab[wib];
  const int i0 = 1 + tx * TX, j0 = 1 + ty * MRY;
  const long qoff = (long)(q + 1) * L.sz, cq = (long)(q + 1) * csz;
  const double lam = alpha + __ldg(lamq + q + 1);
  MCols C;
  march_cols<H>(L, i0, lam, lane, C);
  march_rows<H>(L, j0, lane, rt);
  const bool cin0 = 2 * lane >= H && 2 * lane < PW - H && (i0 + 2 * lane - H) <= L.n1;
  const bool cin1 = 2 * lane + 1 >= H && 2 * lane + 1 < PW - H && (i0 + 2 * lane + 1 - H) <= L.n1;
  const int gu0 = i0 + 2 * lane - H;
  const int I0 = L.c1 ? (C.gi0 + 1) >> 1 : C.gi0, I1 = L.c1 ? (C.gi1 + 1) >> 1 : C.gi1;
  double xa[6], xb[6], ba[6], bb[6];

#define MARCH_LOAD(t, S)                                                                           \
  {                                                                                                 \
    const int gj_ = (int)rt[5][t];                                                                  \
    double v0_ = 0.0, v1_ = 0.0, b0_ = 0.0, b1_ = 0.0;                                              \
    if (gj_ > 0) {                                                                                  \
      const long o_ = qoff + L.sy * gj_;                                                            \
      const long c_ = cq + csy * (L.c2 ? (gj_ + 1) >> 1 : gj_);                                     \
      if (C.gi0 > 0) { v0_ = xin[o_ + C.gi0] + __ldg(cx + c_ + I0); b0_ = b[o_ + C.gi0]; }          \
      if (C.gi1 > 0) { v1_ = xin[o_ + C.gi1] + __ldg(cx + c_ + I1); b1_ = b[o_ + C.gi1]; }          \
    }                                                                                               \
    const double s_ = rt[4][t];                                                                     \
    xa[S] = v0_; xb[S] = v1_; ba[S] = b0_ * (C.rw0 * s_); bb[S] = b1_ * (C.rw1 * s_);               \
  }

#define MARCH_UP_STEP(K)                                                                            \
  if (t0 + (K) < HT + 0) {                                                                          \
    const int t = t0 + (K);                                                                         \
    if (t + 1 < HT) MARCH_LOAD(t + 1, ((K) + 1) % 6)                                                \
    if (t >= 2) { /* colour 1 of row t-1: column 2*lane+1 in even rows */                           \
      constexpr int R = ((K) + 5) % 6, Rm = ((K) + 4) % 6, Rp = (K);                                \
      if ((((K) + 1) & 1) == 0) MARCH_UPD(0, xb[R], xa[R], xb[Rm], xb[Rp], bb[R], t - 1)            \
      else MARCH_UPD(1, xa[R], xb[R], xa[Rm], xa[Rp], ba[R], t - 1)                                 \
    }                                                                                               \
    if (t >= 4) { /* colour 0 of row t-2: column 2*lane in even rows; then the row is final */      \
      constexpr int R = ((K) + 4) % 6, Rm = ((K) + 3) % 6, Rp = ((K) + 5) % 6;                      \
      if (((K) & 1) == 0) MARCH_UPD(1, xa[R], xb[R], xa[Rm], xa[Rp], ba[R], t - 2)                  \
      else MARCH_UPD(0, xb[R], xa[R], xb[Rm], xb[Rp], bb[R], t - 2)                                 \
      const int r = t - 2, gv = j0 + r - H;                                                         \
      if (r < H + MRY && gv <= L.n2) {                                                              \
        const long o = qoff + L.sy * gv;                                                            \
        if (cin0) xout[o + gu0] = xa[R];                                                            \
        if (cin1) xout[o + gu0 + 1] = xb[R];                                                        \
      }                                                                                             \
    }                                                                                               \
  }

  MARCH_LOAD(0, 0)
  for (int t0 = 0; t0 < HT; t0 += 6) {
    MARCH_UP_STEP(0) MARCH_UP_STEP(1) MARCH_UP_STEP(2) MARCH_UP_STEP(3) MARCH_UP_STEP(4) MARCH_UP_STEP(5)
  }
#undef MARCH_UP_STEP
#undef MARCH_LOAD
}
#undef MARCH_UPD

// ---- whole coarse part of the V-cycle in shared memory: one CTA per wavenumber ----------------
constexpr int PMAXLEV = 16;
struct PCoarseArgs {
  PLevel lev[PMAXLEV];
  int nlev;
  const double *b;       // right-hand side of the top level
  const double *xin;     // initial iterate of the top level, or nullptr (zero)
  double *xout;          // result on the top level
  const double *lamq;
  double alpha;
  double *norms;         // only when the top level is level 0
  int nq, want_b;
  int npre, npost, nbot;
  const int *active;
};

__device__ inline void pc_sweep(const PLevel &L, double *x, const double *b, int color, double lam) {
  const int n1 = L.n1, n2 = L.n2, h1 = (n1 + 1) >> 1;
  for (int t = threadIdx.x; t < h1 * n2; t += blockDim.x) {   // only the cells of this colour
    const int j = t / h1 + 1;
    const int i = 2 * (t % h1) + 2 - ((color + j) & 1);
    if (i > n1) continue;
    const double w1 = __ldg(L.w1 + i), w2 = __ldg(L.w2 + j);
    const double diag = w2 * __ldg(L.ad1 + i) + w1 * __ldg(L.ad2 + j) + lam * w1 * w2;
    if (diag == 0.0) continue;
    const int im = i > 1 ? i - 1 : (L.per1 ? n1 : 1), ip = i < n1 ? i + 1 : (L.per1 ? 1 : n1);
    const int jm = j > 1 ? j - 1 : (L.per2 ? n2 : 1), jp = j < n2 ? j + 1 : (L.per2 ? 1 : n2);
    const int c = (i - 1) + n1 * (j - 1);
    const double off = w2 * (__ldg(L.am1 + i) * x[(im - 1) + n1 * (j - 1)] + __ldg(L.ap1 + i) * x[(ip - 1) + n1 * (j - 1)]) +
                       w1 * (__ldg(L.am2 + j) * x[(i - 1) + n1 * (jm - 1)] + __ldg(L.ap2 + j) * x[(i - 1) + n1 * (jp - 1)]);
    x[c] = (b[c] - off) * __drcp_rn(diag);
  }
  __syncthreads();
}

__device__ inline double pc_resid(const PLevel &L, const double *x, const double *b, int i, int j, double lam) {
  const int n1 = L.n1, n2 = L.n2;
  const double w1 = __ldg(L.w1 + i), w2 = __ldg(L.w2 + j);
  const double diag = w2 * __ldg(L.ad1 + i) + w1 * __ldg(L.ad2 + j) + lam * w1 * w2;
  const int im = i > 1 ? i - 1 : (L.per1 ? n1 : 1), ip = i < n1 ? i + 1 : (L.per1 ? 1 : n1);
  const int jm = j > 1 ? j - 1 : (L.per2 ? n2 : 1), jp = j < n2 ? j + 1 : (L.per2 ? 1 : n2);
  const double off = w2 * (__ldg(L.am1 + i) * x[(im - 1) + n1 * (j - 1)] + __ldg(L.ap1 + i) * x[(ip - 1) + n1 * (j - 1)]) +
                     w1 * (__ldg(L.am2 + j) * x[(i - 1) + n1 * (jm - 1)] + __ldg(L.ap2 + j) * x[(i - 1) + n1 * (jp - 1)]);
  return b[(i - 1) + n1 * (j - 1)] - (diag * x[(i - 1) + n1 * (j - 1)] + off);
}

__global__ void __launch_bounds__(512) mg_plane_coarse_kernel(const PCoarseArgs a) {
  extern __shared__ double sm[];
  const int q = blockIdx.x;
  if (a.active && !a.active[q]) return;
  const double lam = a.alpha + __ldg(a.lamq + q + 1);
  // shared-memory layout: x_l, b_l for every level
  double *xs[PMAXLEV], *bs[PMAXLEV];
  {
    double *p = sm;
    for (int l = 0; l < a.nlev; ++l) {
      const int sz = a.lev[l].n1 * a.lev[l].n2;
      xs[l] = p; p += sz;
      bs[l] = p; p += sz;
    }
  }
  {
    const PLevel &T = a.lev[0];
    const long qoff = (long)(q + 1) * T.sz;
    for (int t = threadIdx.x; t < T.n1 * T.n2; t += blockDim.x) {
      const long o = qoff + (t % T.n1 + 1) + T.sy * (t / T.n1 + 1);
      bs[0][t] = a.b[o];
      xs[0][t] = a.xin ? a.xin[o] : 0.0;
    }
    __syncthreads();
  }
  double r2 = 0.0, b2 = 0.0;
  for (int l = 0; l < a.nlev - 1; ++l) {
    const PLevel &F = a.lev[l];
    for (int s = 0; s < a.npre; ++s) { pc_sweep(F, xs[l], bs[l], 0, lam); pc_sweep(F, xs[l], bs[l], 1, lam); }
    const PLevel &C = a.lev[l + 1];
    for (int t = threadIdx.x; t < C.n1 * C.n2; t += blockDim.x) {
      const int I = t % C.n1 + 1, J = t / C.n1 + 1;
      const int i0 = F.c1 ? 2 * I - 1 : I, i1 = F.c1 ? min(2 * I, F.n1) : I;
      const int j0 = F.c2 ? 2 * J - 1 : J, j1 = F.c2 ? min(2 * J, F.n2) : J;
      double s = 0.0;
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) {
          const double r = pc_resid(F, xs[l], bs[l], i, j, lam);
          s += r;
          if (l == 0 && a.norms) {
            r2 += r * r;
            const double bb = bs[0][(i - 1) + F.n1 * (j - 1)];
            b2 += bb * bb;
          }
        }
      bs[l + 1][t] = s;
      xs[l + 1][t] = 0.0;
    }
    __syncthreads();
  }
  {
    const int l = a.nlev - 1;
    const PLevel &B = a.lev[l];
    const int ns = min(a.nbot, B.n1 * B.n2);   // a 1 x 1 level is solved by one update
    for (int s = 0; s < ns; ++s) { pc_sweep(B, xs[l], bs[l], 0, lam); pc_sweep(B, xs[l], bs[l], 1, lam); }
    if (a.nlev == 1 && a.norms) {
      for (int t = threadIdx.x; t < B.n1 * B.n2; t += blockDim.x) {
        const double r = pc_resid(B, xs[l], bs[l], t % B.n1 + 1, t / B.n1 + 1, lam);
        r2 += r * r;
        b2 += bs[l][t] * bs[l][t];
      }
    }
  }
  for (int l = a.nlev - 2; l >= 0; --l) {
    const PLevel &F = a.lev[l];
    const PLevel &C = a.lev[l + 1];
    for (int t = threadIdx.x; t < F.n1 * F.n2; t += blockDim.x) {
      const int i = t % F.n1 + 1, j = t / F.n1 + 1;
      const int I = F.c1 ? (i + 1) >> 1 : i, J = F.c2 ? (j + 1) >> 1 : j;
      xs[l][t] += xs[l + 1][(I - 1) + C.n1 * (J - 1)];
    }
    __syncthreads();
    for (int s = 0; s < a.npost; ++s) { pc_sweep(F, xs[l], bs[l], 1, lam); pc_sweep(F, xs[l], bs[l], 0, lam); }
  }
  {
    const PLevel &T = a.lev[0];
    const long qoff = (long)(q + 1) * T.sz;
    for (int t = threadIdx.x; t < T.n1 * T.n2; t += blockDim.x)
      a.xout[qoff + (t % T.n1 + 1) + T.sy * (t / T.n1 + 1)] = xs[0][t];
  }
  if (a.norms) {
    for (int o = 16; o > 0; o >>= 1) {
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    __shared__ double red[2][16];
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = r2; red[1][threadIdx.x >> 5] = b2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { r2 += red[0][w]; b2 += red[1][w]; }
      atomicAdd(a.norms + q, r2);
      if (a.want_b) atomicAdd(a.norms + a.nq + q, b2);
    }
  }
}

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
  if (const char *e = getenv("SNAC_MG_NSWEEP")) def = std::max(1, atoi(e));
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

void MGSolver::setup(Ctx &ctx, std::vector<SBox> &&boxes_, bool batched, int max_levels) {
  sboxes = std::move(boxes_);
  fft_dir = batched ? 2 : -1;
  ids_local = batched;
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
      c.x.alloc(c.g); c.b.alloc(c.g);
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
  // halo exchange of x on level 0 (coarser ones were built above)
  {
    std::vector<ExField> ef;
    for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, levels[0]->boxes[lb].xp, levels[0]->boxes[lb].g});
    levels[0]->ex.build(ctx, ef, &nbt, nullptr, ids_local);
  }
  // ---- plane-local fast path -----------------------------------------------------------------
  plane_local = batched && !ctx.prm.mg_generic;
  {  // A/B switch between the three implementations of the fused fine-level kernels
    const char *k = getenv("SNAC_MG_KERNEL");
    plane_kernel = !k ? 0 : (!strcmp(k, "tile") ? 1 : (!strcmp(k, "march") ? 2 : 0));
    use_tiles = plane_kernel == 1;
  }
  int npre, npost;
  mg_sweeps(ctx, *this, npre, npost);
  if (npre != npost || npre > 2) plane_local = false;
  if (npre == 2 && plane_kernel != 0) plane_local = false;   // only the prefetch-ring kernels do two sweeps per leg
  for (int lb = 0; lb < nl && plane_local; ++lb)
    for (int d = 0; d < 2; ++d)
      for (int ib = 0; ib < 2; ++ib) {
        const int nb = sboxes[lb].nb[d][ib];
        if (nb >= 0 && (ids_local ? nb : ctx.local_of_rank[nb]) != lb) plane_local = false;   // the plane is split over boxes
      }
  first_smem_level = 0;
  coarse_old = getenv("SNAC_MG_COARSE") && !strcmp(getenv("SNAC_MG_COARSE"), "old");
  // largest plane handed to the previous one-CTA-per-plane kernel (it is latency-bound: keep its top level small)
  const size_t smem_cells = getenv("SNAC_MG_SMEM_CELLS") ? (size_t)atol(getenv("SNAC_MG_SMEM_CELLS")) : 2048;
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
      if (coarse_old) {
        if (nlev - l > PMAXLEV || worst * 2 * sizeof(double) > 160 * 1024 || (l < nlev - 1 && top > smem_cells)) ok = false;
      } else {
        const size_t elems = 2 * worst + 5 * tabs;
        const size_t bytes = elems * (l == 0 ? sizeof(double) : sizeof(float));
        if (nlev - l > CMAXLEV || bytes > 200 * 1024) ok = false;
        if (getenv("SNAC_MG_SMEM_CELLS") && l < nlev - 1 && top > smem_cells) ok = false;   // test hook: force the streaming kernels
      }
      if (!ok) break;
      ls = l;
    }
    coarse_single = !coarse_old && ls > 0;
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
        SNAC_CUDA(cudaMemsetAsync(c.x.buf.raw, 0, c.x.buf.elems * sizeof(double), ctx.stream));
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
  PLevel plevel(int l, int lb) const {
    const MGLevelBox &B = S.levels[l]->boxes[lb];
    PLevel P;
    P.n1 = B.g.n[0]; P.n2 = B.g.n[1];
    P.sy = B.g.sy; P.sz = B.g.sz;
    P.am1 = B.am[0].p(); P.ap1 = B.ap[0].p(); P.ad1 = B.ad[0].p(); P.w1 = B.w[0].p();
    P.am2 = B.am[1].p(); P.ap2 = B.ap[1].p(); P.ad2 = B.ad[1].p(); P.w2 = B.w[1].p();
    P.sam1 = B.sam[0].p(); P.sap1 = B.sap[0].p(); P.se1 = B.se[0].p(); P.srw1 = B.srw[0].p();
    P.sam2 = B.sam[1].p(); P.sap2 = B.sap[1].p(); P.se2 = B.se[1].p(); P.srw2 = B.srw[1].p();
    P.c1 = S.levels[l]->coarsen[0]; P.c2 = S.levels[l]->coarsen[1];
    P.per1 = S.sboxes[lb].nb[0][0] >= 0; P.per2 = S.sboxes[lb].nb[1][0] >= 0;
    return P;
  }
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
      PLevel P = plevel(l, (int)lb);
      MGLevelBox &c = S.levels[l + 1]->boxes[lb];
      const double *xi = l == 0 ? xbuf(l, (int)lb, 0) : nullptr;   // below level 0 the iterate is zero on entry
      if (S.plane_kernel == 0) {
        launch_ring_down(ctx.stream, rlevel(l, (int)lb), npre, xi, xbuf(l, (int)lb, 1), F.boxes[lb].bp, c.bp, c.g.sy, c.g.sz, S.lam[lb].p(),
                         S.alpha, norms ? S.norms.raw : nullptr, S.nq, want_b ? 1 : 0, S.d_active, S.est_active, (l == 0 && want_b) ? S.guess_scale : 1.0);
      } else if (S.use_tiles) {
        dim3 grd((P.n1 + ptx(PH_D) - 1) / ptx(PH_D), (P.n2 + PT_Y - 1) / PT_Y, S.nq);
        mg_plane_down_kernel<<<grd, 256, 0, ctx.stream>>>(P, xi, xbuf(l, (int)lb, 1), F.boxes[lb].bp, c.bp, c.g.sy, c.g.sz,
                                                          S.lam[lb].p(), S.alpha, norms ? S.norms.raw : nullptr, S.nq, want_b ? 1 : 0, S.d_active);
      } else {
        const int ntx = (P.n1 + ptx(3) - 1) / ptx(3), nty = (P.n2 + MRY - 1) / MRY;
        const long nw = (long)ntx * nty * S.nq;
        mg_march_down_kernel<<<(unsigned)((nw + 7) / 8), 256, 0, ctx.stream>>>(P, xi, xbuf(l, (int)lb, 1), F.boxes[lb].bp, c.bp, c.g.sy,
                                                                              c.g.sz, S.lam[lb].p(), S.alpha,
                                                                              norms ? S.norms.raw : nullptr, S.nq, want_b ? 1 : 0, ntx, nty, S.d_active);
      }
      ctx.launches++;
    }
  }
  void plane_up(int l) {
    ProfScope ps(ctx, l == 0 ? PC_MG_UP0 : PC_MG_COARSE, l == 0 ? active_points(0) : 0.0);
    MGLevel &F = *S.levels[l];
    for (size_t lb = 0; lb < F.boxes.size(); ++lb) {
      PLevel P = plevel(l, (int)lb);
      MGLevelBox &c = S.levels[l + 1]->boxes[lb];
      if (S.plane_kernel == 0) {
        launch_ring_up(ctx.stream, rlevel(l, (int)lb), npost, xbuf(l, (int)lb, 1), xbuf(l, (int)lb, 0), F.boxes[lb].bp, c.xp, c.g.sy, c.g.sz,
                       S.lam[lb].p(), S.alpha, S.nq, S.d_active, S.est_active);
      } else if (S.use_tiles) {
        dim3 grd((P.n1 + ptx(PH_U) - 1) / ptx(PH_U), (P.n2 + PT_Y - 1) / PT_Y, S.nq);
        mg_plane_up_kernel<<<grd, 256, 0, ctx.stream>>>(P, xbuf(l, (int)lb, 1), xbuf(l, (int)lb, 0), F.boxes[lb].bp, c.xp, c.g.sy,
                                                        c.g.sz, S.lam[lb].p(), S.alpha, S.d_active);
      } else {
        const int ntx = (P.n1 + ptx(2) - 1) / ptx(2), nty = (P.n2 + MRY - 1) / MRY;
        const long nw = (long)ntx * nty * S.nq;
        mg_march_up_kernel<<<(unsigned)((nw + 7) / 8), 256, 0, ctx.stream>>>(P, xbuf(l, (int)lb, 1), xbuf(l, (int)lb, 0), F.boxes[lb].bp,
                                                                            c.xp, c.g.sy, c.g.sz, S.lam[lb].p(), S.alpha, S.nq, ntx, nty, S.d_active);
      }
      ctx.launches++;
    }
  }
  void plane_coarse(bool norms, bool want_b) {
    const int ls = S.first_smem_level, nlev = (int)S.levels.size();
    ProfScope ps(ctx, ls == 0 ? PC_MG_DOWN0 : PC_MG_COARSE);
    for (size_t lb = 0; lb < S.levels[ls]->boxes.size() && !S.coarse_old; ++lb) {
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
      {
        static const bool adapt = !(getenv("SNAC_MG_COARSE_ADAPT") && atoi(getenv("SNAC_MG_COARSE_ADAPT")) == 0);
        a.adaptive = (ls > 0 && adapt) ? 1 : 0;
      }
      if (ls > 0) {   // tuning hooks for the FP32 coarse-grid correction: sweeps per leg, depth of the hierarchy
        static const int csw = getenv("SNAC_MG_COARSE_NSW") ? atoi(getenv("SNAC_MG_COARSE_NSW")) : 0;
        static const int cml = getenv("SNAC_MG_COARSE_MAXLEV") ? atoi(getenv("SNAC_MG_COARSE_MAXLEV")) : 0;
        static const int cnb = getenv("SNAC_MG_COARSE_NBOT") ? atoi(getenv("SNAC_MG_COARSE_NBOT")) : 0;
        if (csw > 0) a.npre = a.npost = a.npre_low = a.npost_low = csw;
        static const int clo = getenv("SNAC_MG_COARSE_NSW_LOW") ? atoi(getenv("SNAC_MG_COARSE_NSW_LOW")) : 0;
        if (clo > 0) a.npre_low = a.npost_low = clo;
        if (cml > 0) a.nlev = std::min(a.nlev, cml);
        if (cnb > 0) a.nbot = cnb;
      }
      a.active = S.d_active;
      launch_mg_coarse(ctx.stream, a, S.coarse_single);
      ctx.launches++;
    }
    for (size_t lb = 0; lb < S.levels[ls]->boxes.size() && S.coarse_old; ++lb) {
      PCoarseArgs a;
      a.nlev = nlev - ls;
      size_t cells = 0;
      for (int l = ls; l < nlev; ++l) {
        a.lev[l - ls] = plevel(l, (int)lb);
        cells += (size_t)a.lev[l - ls].n1 * a.lev[l - ls].n2;
      }
      MGLevelBox &T = S.levels[ls]->boxes[lb];
      a.b = T.bp;
      a.xin = ls == 0 ? T.xp : nullptr;
      a.xout = T.xp;
      a.lamq = S.lam[lb].p();
      a.alpha = S.alpha;
      a.norms = (ls == 0 && norms) ? S.norms.raw : nullptr;
      a.nq = S.nq; a.want_b = want_b ? 1 : 0;
      a.npre = npre; a.npost = npost; a.nbot = nbot;
      a.active = S.d_active;
      const size_t smem = cells * 2 * sizeof(double);
      if (smem > 48 * 1024)
        SNAC_CUDA(cudaFuncSetAttribute(mg_plane_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      mg_plane_coarse_kernel<<<S.nq, 512, smem, ctx.stream>>>(a);
      ctx.launches++;
    }
  }
};
}  // namespace

void MGSolver::stats(Ctx &ctx, int *iters, double *relres, double *mean_iters) {
  if (plane_local) {
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
    if (getenv("SNAC_MG_DUMP")) {   // diagnostic: V-cycles per wavenumber
      fprintf(stderr, "mg iters per plane:");
      for (int q = 0; q < nq; ++q) fprintf(stderr, " %d", it[q]);
      fprintf(stderr, "\n");
    }
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
  static const bool use_hist = !(getenv("SNAC_MG_HISTORY") && atoi(getenv("SNAC_MG_HISTORY")) == 0);
  const int key = S.hist_key & (MGSolver::HIST_KEYS - 1);
  const int *prev = S.hist(key, 1 - S.hist_cur[key]);
  int *cur = S.hist(key, S.hist_cur[key]);
  const int prev_n = use_hist ? S.hist_n[key] : 0;
  int polled = nq;
  auto predict = [&](int i) -> int {   // planes still iterating after cycle i
    if (i < 0) return nq;
    if (i < prev_n) return std::min(nq, std::max(1, prev[i] + prev[i] / 8 + 1));
    return std::max(1, polled);
  };
  {   // initial guess: the last solution of this key (see MGSolver::xslot)
    static const bool per_key = !(getenv("SNAC_MG_GUESS_PER_KEY") && atoi(getenv("SNAC_MG_GUESS_PER_KEY")) == 0);
    if (per_key && in_alt && key != S.cur_key && S.slot_key[key] >= 0) {
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
    static const bool no_swap = getenv("SNAC_MG_SWAP") && atoi(getenv("SNAC_MG_SWAP")) == 0;
    const bool swap = in_alt && !no_swap;
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
  int npre, npost;
  mg_sweeps(ctx, *this, npre, npost);
  Cycle cy{ctx, *this, {fft_dir != 0, fft_dir != 1, fft_dir != 2}, npre, npost,
           ctx.prm.mg_nbottom > 0 ? ctx.prm.mg_nbottom : 20};
  if (plane_local) {
    solve_plane_local(ctx, *this, cy);
    return;
  }
  // general multi-box path: the host reads the per-plane residual norms once per cycle
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
    SNAC_CUDA(cudaMemcpyAsync(h_norms.data(), norms.raw, nn, cudaMemcpyDeviceToHost, ctx.stream));
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
