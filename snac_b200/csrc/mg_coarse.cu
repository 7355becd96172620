// K4, plane-local fast path: every level of a V-cycle from level `ls` down in ONE kernel, one CTA
// per wavenumber plane, all iterates and right-hand sides of those levels resident in shared
// memory (see mg.cu for the operator, the cycle and the reference call sites,
// src/solver.f90:843-922).
//
// When the kernel starts at level 0 (planes small enough to fit) it is the whole solver cycle and
// works in FP64.  When it starts below level 0 it only produces the coarse-grid correction of the
// FP64 fine-level kernels (mg_ring.cu): that correction is computed in FP32 -- an error of 1e-7
// relative to a correction that itself only has to reduce the fine residual by a factor of ~4 per
// cycle does not change the convergence, the fine-level residual and the iterate stay FP64 -- which
// halves the shared-memory footprint, so that twice as large a level (e.g. 128 x 128 for 256 x 256
// planes) can be taken off the HBM-streaming kernels.
//
// Equations are used in the scaled form (division by w1*w2, as in mg_ring.cu):
//   x = ( b~ + na1m xL + na1p xR + na2m xD + na2p xU ) / (e1 + e2 + lambda)
// and the restriction of the (unscaled) residual by summation over the children becomes
//   b~_c(I,J) = sum_children (w1_i / W1_I)(w2_j / W2_J) r~(i,j),   W = width of the coarse cell.
#include "common.h"

#ifndef SNAC_COARSE_WARP_CELLS
#define SNAC_COARSE_WARP_CELLS 256   // levels with at most this many cells are run by warp 0 alone
#endif

namespace snac {

namespace {

struct CLev {      // one level as laid out in shared memory (offsets in elements from the base)
  int n1, n2, c1, c2, per1, per2;
  int ox, ob;      // x, b: [n2][n1]
  int o1, o2;      // tables of directions 0/1: [5][n]: na_m, na_p, e, restriction weight w/W, w
};

__device__ __forceinline__ double c_rcp(double e) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(e));
  double t = fma(-e, r, 1.0);
  r = fma(r, t, r);
  t = fma(-e, r, 1.0);
  return fma(r, t, r);
}
__device__ __forceinline__ float c_rcp(float e) {   // 1 ulp is plenty for a coarse-grid smoother
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e));
  return r;
}

// A phase is executed by a group of nthr threads (the whole CTA, or warp 0 alone on the smallest
// levels, where a block-wide barrier per phase would cost more than the work).  2-D thread layout
// of a level inside the group: W = 2^lw threads along direction 0, the rest along direction 1.
struct CGroup {
  int tid, nthr;   // nthr is a power of two
  bool warp;       // the group is one warp: phases are separated by __syncwarp
  __device__ __forceinline__ void sync() const {
    if (warp) __syncwarp();
    else __syncthreads();
  }
  __device__ __forceinline__ int lw(int n) const {
    const int m = min(max(n, 1), nthr);
    return 32 - __clz(m - 1);   // ceil(log2(m))
  }
};

// residual of the scaled equation at cell (i, j)
template <class T>
__device__ __forceinline__ T c_resid(const T *sm, const CLev &L, int i, int j, T lam) {
  const int n1 = L.n1, n2 = L.n2;
  const int im = i > 0 ? i - 1 : (L.per1 ? n1 - 1 : 0), ip = i < n1 - 1 ? i + 1 : (L.per1 ? 0 : n1 - 1);
  const int jm = j > 0 ? j - 1 : (L.per2 ? n2 - 1 : 0), jp = j < n2 - 1 ? j + 1 : (L.per2 ? 0 : n2 - 1);
  const T *x = sm + L.ox, *t1 = sm + L.o1, *t2 = sm + L.o2;
  const T e = t1[2 * n1 + i] + t2[2 * n2 + j] + lam;
  T acc = sm[L.ob + j * n1 + i];
  acc = fma(t1[i], x[j * n1 + im], acc);
  acc = fma(t1[n1 + i], x[j * n1 + ip], acc);
  acc = fma(t2[j], x[jm * n1 + i], acc);
  acc = fma(t2[n2 + j], x[jp * n1 + i], acc);
  return fma(-e, x[j * n1 + i], acc);
}

// one colour of a Gauss-Seidel sweep: a thread keeps its column and walks over rows of one parity
#ifndef SNAC_SWEEP_PAIRS
// One colour of a Gauss-Seidel sweep.  Consecutive lanes take consecutive columns and zigzag over a
// pair of rows (column i updates row 2 jp + ((i + colour) & 1)), so that the 32 accesses of a warp
// fall into 32 different banks whenever the row length is a multiple of 32 (the column-pair layout
// kept below for comparison has every lane skip a column: two-way conflicts on every access to x
// and b; measured 10-30 % slower per level).
template <class T>
__device__ __forceinline__ void c_sweep(T *sm, const CLev &L, int color, T lam, const CGroup &G) {
  const int n1 = L.n1, n2 = L.n2, per1 = L.per1, per2 = L.per2;
  const int lw = G.lw(n1), W = 1 << lw, tx = G.tid & (W - 1), ty = G.tid >> lw, nty = G.nthr >> lw;
  T *x = sm + L.ox;
  const T *b = sm + L.ob, *t1 = sm + L.o1, *t2 = sm + L.o2;
  if (G.tid < G.nthr) {
    const int np2 = (n2 + 1) >> 1;
    const int wlo = per2 ? (n2 - 1) * n1 : 0, whi = per2 ? -(n2 - 1) * n1 : 0;   // row offsets of the wrapped / clamped neighbours
    for (int i = tx; i < n1; i += W) {
      const int z = (i + color) & 1;
      const int im = i > 0 ? i - 1 : (per1 ? n1 - 1 : 0), ip = i < n1 - 1 ? i + 1 : (per1 ? 0 : n1 - 1);
      const T cm = t1[i], cp = t1[n1 + i], e1 = t1[2 * n1 + i] + lam;
      T *xc = x + i;
      const T *xl = x + im, *xr = x + ip, *bc = b + i;
#pragma unroll 2
      for (int jp = ty; jp < np2; jp += nty) {
        const int j = 2 * jp + z;
        if (j < n2) {
          const int o = j * n1;
          const int om = j > 0 ? o - n1 : o + wlo, op = j < n2 - 1 ? o + n1 : o + whi;
          const T e = e1 + t2[2 * n2 + j];
          T acc = bc[o];
          T acc2 = t2[j] * xc[om];
          acc = fma(cm, xl[o], acc);
          acc2 = fma(t2[n2 + j], xc[op], acc2);
          acc = fma(cp, xr[o], acc);
          // e == 0: a singular one-cell level, nothing to solve (the update would be 0/0)
          if (e != (T)0) xc[o] = (acc + acc2) * c_rcp(e);
        }
      }
    }
  }
  G.sync();
}
#else
template <class T>
__device__ __forceinline__ void c_sweep(T *sm, const CLev &L, int color, T lam, const CGroup &G) {
  const int n1 = L.n1, n2 = L.n2, h1 = (n1 + 1) >> 1, per1 = L.per1, per2 = L.per2;
  const int lw = G.lw(h1), W = 1 << lw, tx = G.tid & (W - 1), ty = G.tid >> lw, nty = G.nthr >> lw;
  T *x = sm + L.ox;
  const T *b = sm + L.ob, *t1 = sm + L.o1, *t2 = sm + L.o2;
  if (G.tid < G.nthr) {
    // rows j = ty, ty + nty, ...: all of one parity when nty is even, so the column is fixed;
    // with a single row of threads (nty == 1) walk the two parities one after the other
    const int npar = (nty & 1) ? 2 : 1, jstep = (nty & 1) ? 2 * nty : nty;
    const int wlo = per2 ? (n2 - 1) * n1 : 0, whi = per2 ? -(n2 - 1) * n1 : 0;   // row offsets of the wrapped / clamped neighbours
    for (int p = tx; p < h1; p += W) {
      for (int par = 0; par < npar; ++par) {
        const int jfirst = ty + par * nty;
        const int i = 2 * p + ((color + jfirst) & 1);
        if (i >= n1) continue;
        const int im = i > 0 ? i - 1 : (per1 ? n1 - 1 : 0), ip = i < n1 - 1 ? i + 1 : (per1 ? 0 : n1 - 1);
        const T cm = t1[i], cp = t1[n1 + i], e1 = t1[2 * n1 + i] + lam;
        T *xc = x + i;
        const T *xl = x + im, *xr = x + ip, *bc = b + i;
#ifdef SNAC_SWEEP_PEEL
#define SNAC_CELL(J, DM, DP)                                                        \
  {                                                                                 \
    const int o = (J) * n1;                                                         \
    const T e = e1 + t2[2 * n2 + (J)];                                              \
    T acc = bc[o];                                                                  \
    T acc2 = t2[(J)] * xc[o + (DM)];                                                \
    acc = fma(cm, xl[o], acc);                                                      \
    acc2 = fma(t2[n2 + (J)], xc[o + (DP)], acc2);                                   \
    acc = fma(cp, xr[o], acc);                                                      \
    if (e != (T)0) xc[o] = (acc + acc2) * c_rcp(e);                                 \
  }
        int j = jfirst;
        if (j == 0) { SNAC_CELL(0, wlo, (n2 > 1 ? n1 : whi)) j += jstep; }
#pragma unroll 2
        for (; j < n2 - 1; j += jstep) SNAC_CELL(j, -n1, n1)
        if (j == n2 - 1) SNAC_CELL(j, (n2 > 1 ? -n1 : wlo), whi)
#undef SNAC_CELL
#else
#pragma unroll 2
        for (int j = jfirst; j < n2; j += jstep) {
          const int o = j * n1;
          const int om = j > 0 ? o - n1 : o + wlo, op = j < n2 - 1 ? o + n1 : o + whi;
          const T e = e1 + t2[2 * n2 + j];
          T acc = bc[o];
          T acc2 = t2[j] * xc[om];
          acc = fma(cm, xl[o], acc);
          acc2 = fma(t2[n2 + j], xc[op], acc2);
          acc = fma(cp, xr[o], acc);
          // e == 0: a singular one-cell level, nothing to solve (the update would be 0/0)
          xc[o] = e != (T)0 ? (acc + acc2) * c_rcp(e) : xc[o];
        }
#endif
      }
    }
  }
  G.sync();
}
#endif

// pre-relaxation on level F, residual restricted to level C (whose iterate is zeroed)
template <class T>
__device__ __forceinline__ void c_down(T *sm, const CLev &F, const CLev &C, T lam, int npre, const CGroup &G, bool want_norm,
                                       double &r2, double &b2) {
  for (int s = 0; s < npre; ++s) { c_sweep(sm, F, 0, lam, G); c_sweep(sm, F, 1, lam, G); }
  const T *f1 = sm + F.o1, *f2 = sm + F.o2;
  if (F.c1 && F.c2 && F.n1 >= 2 && !want_norm) {
    // full coarsening: of the four children of a coarse cell only (2I, 2J) and (2I+1, 2J+1) have a
    // residual after the second colour.  One lane per fine column (conflict-free, as in the sweeps),
    // row 2J + (i & 1); the two children are neighbouring lanes and are added by a shuffle.
    const int lwf = G.lw(F.n1), Wf = 1 << lwf, txf = G.tid & (Wf - 1), tyf = G.tid >> lwf, ntyf = G.nthr >> lwf;
    if (G.tid < G.nthr) {
      for (int ib = 0; ib < F.n1; ib += Wf) {
        const int i = ib + txf;
        for (int Jb = 0; Jb < C.n2; Jb += ntyf) {
          const int J = Jb + tyf, j = 2 * J + (i & 1);
          T v = (T)0;
          if (i < F.n1 && J < C.n2 && j < F.n2) v = f1[3 * F.n1 + i] * f2[3 * F.n2 + j] * c_resid(sm, F, i, j, lam);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          if (!(i & 1) && i < F.n1 && J < C.n2) {
            sm[C.ob + J * C.n1 + (i >> 1)] = v;
            sm[C.ox + J * C.n1 + (i >> 1)] = (T)0;
          }
        }
      }
    }
    G.sync();
    return;
  }
  const int lw = G.lw(C.n1), W = 1 << lw, tx = G.tid & (W - 1), ty = G.tid >> lw, nty = G.nthr >> lw;
  if (G.tid < G.nthr) {
    for (int I = tx; I < C.n1; I += W) {
      const int i0 = F.c1 ? 2 * I : I, i1 = F.c1 ? min(2 * I + 1, F.n1 - 1) : I;
      for (int J = ty; J < C.n2; J += nty) {
        const int j0 = F.c2 ? 2 * J : J, j1 = F.c2 ? min(2 * J + 1, F.n2 - 1) : J;
        T sum = (T)0;
        for (int j = j0; j <= j1; ++j)
          for (int i = i0; i <= i1; ++i) {
            if (want_norm) {
              const double bb = (double)sm[F.ob + j * F.n1 + i] * ((double)f1[4 * F.n1 + i] * (double)f2[4 * F.n2 + j]);
              b2 += bb * bb;
            }
            if ((i + j) & 1) continue;   // after the second colour its own residual vanishes
            const T r = c_resid(sm, F, i, j, lam);
            sum = fma(f1[3 * F.n1 + i] * f2[3 * F.n2 + j], r, sum);
            if (want_norm) {
              const double rr = (double)r * ((double)f1[4 * F.n1 + i] * (double)f2[4 * F.n2 + j]);
              r2 += rr * rr;
            }
          }
        sm[C.ob + J * C.n1 + I] = sum;
        sm[C.ox + J * C.n1 + I] = (T)0;
      }
    }
  }
  G.sync();
}

// correction from level C added to level F, post-relaxation
template <class T>
__device__ __forceinline__ void c_up(T *sm, const CLev &F, const CLev &C, T lam, int npost, const CGroup &G) {
  const int lw = G.lw(F.n1), W = 1 << lw, tx = G.tid & (W - 1), ty = G.tid >> lw, nty = G.nthr >> lw;
  if (G.tid < G.nthr) {
    for (int i = tx; i < F.n1; i += W) {
      const int I = F.c1 ? i >> 1 : i;
#pragma unroll 4
      for (int j = ty; j < F.n2; j += nty) sm[F.ox + j * F.n1 + i] += sm[C.ox + (F.c2 ? j >> 1 : j) * C.n1 + I];
    }
  }
  G.sync();
  for (int s = 0; s < npost; ++s) { c_sweep(sm, F, 1, lam, G); c_sweep(sm, F, 0, lam, G); }
}

}  // namespace

template <class T>
__global__ void __launch_bounds__(1024) mg_coarse_kernel(const CoarseArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ CLev lv[CMAXLEV];
  __shared__ int toff[2 * CMAXLEV + 1];   // prefix sums of the table lengths over (level, direction)
  __shared__ double red[2][32];
  T *sm = reinterpret_cast<T *>(smem_raw);
  const int q = blockIdx.x;
  if (a.active && !a.active[q]) return;
#ifdef SNAC_COARSE_TIMING
  __shared__ long long tstamp[48];
  int nts = 0;
#define TS() do { __syncthreads(); if (threadIdx.x == 0 && nts < 48) tstamp[nts] = clock64(); ++nts; } while (0)
#else
#define TS() do {} while (0)
#endif
  TS();
  const T lam = (T)(a.alpha + __ldg(a.lamq + q + 1));
  int nlev = a.nlev;
  __shared__ int s_nlev, s_nbot;
  const CGroup GB{(int)threadIdx.x, (int)blockDim.x, false}, GW{(int)threadIdx.x, 32, true};
  if (threadIdx.x == 0) {   // carve the shared memory
    int p = 0;
    for (int l = 0; l < nlev; ++l) {
      CLev &L = lv[l];
      L.n1 = a.lev[l].n1; L.n2 = a.lev[l].n2;
      L.c1 = a.lev[l].c1; L.c2 = a.lev[l].c2; L.per1 = a.lev[l].per1; L.per2 = a.lev[l].per2;
      L.ox = p; p += L.n1 * L.n2;
      L.ob = p; p += L.n1 * L.n2;
      L.o1 = p; p += 5 * L.n1;
      L.o2 = p; p += 5 * L.n2;
    }
    toff[0] = 0;
    for (int l = 0; l < nlev; ++l) {
      toff[2 * l + 1] = toff[2 * l] + lv[l].n1;
      toff[2 * l + 2] = toff[2 * l + 1] + lv[l].n2;
    }
  }
  __syncthreads();
  TS();
  // coefficient tables (global tables are 1-based with 8 doubles per entry: na_m, na_p, e, w, 1/w): one
  // work item per (level, direction, index), all of them in flight at once -- level after level
  // the loop paid one global-memory latency per level and direction, 10 us of an 80 us cycle
  for (int it = threadIdx.x; it < toff[2 * nlev]; it += blockDim.x) {
    int k = 0;
    while (it >= toff[k + 1]) ++k;
    const int l = k >> 1, d = k & 1, i = it - toff[k];
    const CLev &L = lv[l];
    const int n = d ? L.n2 : L.n1;
    const int co = d ? L.c2 : L.c1;
    const double *g = d ? a.lev[l].t2 : a.lev[l].t1;
    const double *gc = l + 1 < nlev ? (d ? a.lev[l + 1].t2 : a.lev[l + 1].t1) : nullptr;
    T *t = sm + (d ? L.o2 : L.o1);
    const double *e = g + 8 * (i + 1);
    const double w = __ldg(e + 3);
    const double wc = (gc && co) ? __ldg(gc + 8 * ((i >> 1) + 1) + 3) : w;
    t[i] = (T)__ldg(e); t[n + i] = (T)__ldg(e + 1); t[2 * n + i] = (T)__ldg(e + 2);
    t[4 * n + i] = (T)w;
    t[3 * n + i] = (gc && co) ? (T)(w / wc) : (T)1;
  }
  // Depth of the hierarchy for THIS wavenumber: on a level where the shift lambda dominates the
  // diagonal, lambda >= r (e1 + e2) with r >= 1/2, Gauss-Seidel itself contracts by about
  // 1/(1+r)^2 per sweep, so that level is solved by a few sweeps and the levels below it (the
  // latency-bound tail of the cycle, run by one warp) are not visited.  Planes with a small shift
  // (the low wavenumbers, lambda -> 0) keep the whole hierarchy.
  TS();
  if (a.adaptive && nlev > 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      int nl = nlev, nb = a.nbot;
      for (int l = 1; l < nlev; ++l) {
        const CLev L = lv[l];
        const double e = (double)sm[L.o1 + 2 * L.n1 + L.n1 / 2] + (double)sm[L.o2 + 2 * L.n2 + L.n2 / 2];
        const double r = e != 0.0 ? (double)lam / e : 0.0;   // > 0: the shift has the sign of the diagonal
        if (r >= 0.5) {
          nl = l + 1;
          nb = min(a.nbot, (int)ceil(2.3 / log1p(r)));   // (1+r)^(-2 nb) <= 1e-2: plenty for a coarse-grid correction
          break;
        }
      }
      s_nlev = nl; s_nbot = nb;
    }
    __syncthreads();
    nlev = s_nlev;
  } else if (threadIdx.x == 0) {
    s_nbot = a.nbot;
  }
  TS();
  // top level: right-hand side (scaled) and iterate; rows of the plane are contiguous in memory
  {
    const CLev L = lv[0];
    const long qoff = (long)(q + 1) * a.sz;
    const int n1 = L.n1, n2 = L.n2;
    const int lw = GB.lw(n1), W = 1 << lw, tx = threadIdx.x & (W - 1), ty = threadIdx.x >> lw, nty = blockDim.x >> lw;
    for (int i = tx; i < n1; i += W) {
      const double s1 = __ldg(a.lev[0].t1 + 8 * (i + 1) + 4);
#pragma unroll 16
      for (int j = ty; j < n2; j += nty) {
        const long o = qoff + (i + 1) + a.sy * (j + 1);
        const double s = s1 * __ldg(a.lev[0].t2 + 8 * (j + 1) + 4);
        sm[L.ob + j * n1 + i] = (T)(__ldg(a.b + o) * s);
        sm[L.ox + j * n1 + i] = a.xin ? (T)(a.xin[o] * a.xscale) : (T)0;
      }
    }
  }
  __syncthreads();
  const int nbot = s_nbot;
  // first level left to warp 0 alone (at most 256 cells); level 0 always belongs to the CTA
  int lsw = nlev;
  while (lsw > 1 && lv[lsw - 1].n1 * lv[lsw - 1].n2 <= SNAC_COARSE_WARP_CELLS) --lsw;
  TS();
  double r2 = 0.0, b2 = 0.0;
  // ---- down leg, CTA-wide levels
  for (int l = 0; l < lsw - 1; ++l) {
    const CLev F = lv[l], C = lv[l + 1];
    c_down(sm, F, C, lam, l == 0 ? a.npre : a.npre_low, GB, l == 0 && a.norms, r2, b2);
    TS();
  }
  // ---- from level lsw-1 down: restriction into the first small level by the CTA, then warp 0 runs
  //      the rest of the cycle on the small levels by itself
  if (lsw < nlev) {
    {
      const CLev F = lv[lsw - 1], C = lv[lsw];
      c_down(sm, F, C, lam, lsw == 1 ? a.npre : a.npre_low, GB, lsw == 1 && a.norms, r2, b2);
    }
    TS();
    if (threadIdx.x < 32) {
      for (int l = lsw; l < nlev - 1; ++l) {
        const CLev F = lv[l], C = lv[l + 1];
        c_down(sm, F, C, lam, a.npre_low, GW, false, r2, b2);
      }
      const CLev B = lv[nlev - 1];
      const int ns = min(nbot, B.n1 * B.n2);   // a 1 x 1 level is solved by one update
      for (int s = 0; s < ns; ++s) { c_sweep(sm, B, 0, lam, GW); c_sweep(sm, B, 1, lam, GW); }
      for (int l = nlev - 2; l >= lsw; --l) {
        const CLev F = lv[l], C = lv[l + 1];
        c_up(sm, F, C, lam, a.npost_low, GW);
      }
    }
    __syncthreads();
    TS();
    {
      const CLev F = lv[lsw - 1], C = lv[lsw];
      c_up(sm, F, C, lam, lsw == 1 ? a.npost : a.npost_low, GB);
    }
    TS();
  } else {   // the coarsest level belongs to the CTA (a single-level or all-large hierarchy)
    const CLev B = lv[nlev - 1];
    const int ns = min(nbot, B.n1 * B.n2);
    for (int s = 0; s < ns; ++s) { c_sweep(sm, B, 0, lam, GB); c_sweep(sm, B, 1, lam, GB); }
    if (nlev == 1 && a.norms) {
      for (int t = threadIdx.x; t < B.n1 * B.n2; t += blockDim.x) {
        const int j = t / B.n1, i = t - j * B.n1;
        const double w = (double)sm[B.o1 + 4 * B.n1 + i] * (double)sm[B.o2 + 4 * B.n2 + j];
        const double rr = (double)c_resid(sm, B, i, j, lam) * w, bb = (double)sm[B.ob + t] * w;
        r2 += rr * rr;
        b2 += bb * bb;
      }
    }
  }
  // ---- up leg, CTA-wide levels
  for (int l = lsw - 2; l >= 0; --l) {
    const CLev F = lv[l], C = lv[l + 1];
    c_up(sm, F, C, lam, l == 0 ? a.npost : a.npost_low, GB);
    TS();
  }
  // ---- result
  {
    const CLev L = lv[0];
    const long qoff = (long)(q + 1) * a.sz;
    const int n1 = L.n1, n2 = L.n2;
    const int lw = GB.lw(n1), W = 1 << lw, tx = threadIdx.x & (W - 1), ty = threadIdx.x >> lw, nty = blockDim.x >> lw;
    for (int i = tx; i < n1; i += W)
#pragma unroll 4
      for (int j = ty; j < n2; j += nty) a.xout[qoff + (i + 1) + a.sy * (j + 1)] = (double)sm[L.ox + j * n1 + i];
  }
  TS();
#ifdef SNAC_COARSE_TIMING
  if (threadIdx.x == 0 && blockIdx.x == 1) {
    printf("coarse q=1 nlev=%d lsw=%d:", nlev, lsw);
    for (int i = 1; i < nts && i < 48; ++i) printf(" %lld", tstamp[i] - tstamp[i - 1]);
    printf("\n");
  }
#endif
  if (a.norms) {
    for (int o = 16; o > 0; o >>= 1) {
      r2 += __shfl_xor_sync(0xffffffffu, r2, o);
      b2 += __shfl_xor_sync(0xffffffffu, b2, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = r2; red[1][threadIdx.x >> 5] = b2; }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { r2 += red[0][w]; b2 += red[1][w]; }
      atomicAdd(a.norms + q, r2);
      if (a.want_b) atomicAdd(a.norms + a.nq + q, b2);
    }
  }
}

size_t mg_coarse_smem(const CoarseArgs &a, bool single) {
  size_t elems = 0;
  for (int l = 0; l < a.nlev; ++l) elems += 2 * (size_t)a.lev[l].n1 * a.lev[l].n2 + 5 * (size_t)(a.lev[l].n1 + a.lev[l].n2);
  return elems * (single ? sizeof(float) : sizeof(double)) + 16;
}

void launch_mg_coarse(cudaStream_t st, const CoarseArgs &a, bool single) {
  const size_t smem = mg_coarse_smem(a, single);
  const long top = (long)a.lev[0].n1 * a.lev[0].n2;
  int nt = 128;
  while (nt < 1024 && nt * 8 < top) nt *= 2;   // about eight cells of the top level per thread
  // opt in as soon as the dynamic part alone comes near 48 KB: the kernel's static shared memory counts against the
  // same default limit (a 64 x 64 hierarchy in FP32 asks for 48 784 B and was refused with "invalid argument")
  static size_t cap[2] = {40 * 1024, 40 * 1024};
  if (smem > cap[single]) {
    if (single) SNAC_CUDA(cudaFuncSetAttribute(mg_coarse_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else SNAC_CUDA(cudaFuncSetAttribute(mg_coarse_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cap[single] = smem;
  }
  if (single) mg_coarse_kernel<float><<<a.nq, nt, smem, st>>>(a);
  else mg_coarse_kernel<double><<<a.nq, nt, smem, st>>>(a);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    throw Error(std::string("mg_coarse_kernel launch failed: ") + cudaGetErrorString(e) + " (planes " + std::to_string(a.nq) + ", threads " +
                std::to_string(nt) + ", dynamic shared memory " + std::to_string(smem) + " B, " + std::to_string(a.nlev) + " levels from " +
                std::to_string(a.lev[0].n1) + " x " + std::to_string(a.lev[0].n2) + ")");
}

}  // namespace snac
