// K3b/K5: batched 1-D real-to-real transforms along one axis of a ghosted box.
//
// Replaces the FFTW guru r2r plans of the reference (fftini, src/fft.f90:10-69; execution
// src/fft.f90:85-94; kind selection find_fft, src/fft.f90:96-149; interface src/fftw.f90:9-54).
// FFTW is an external, unpinned dependency of the reference; the transforms here follow the
// unnormalised definitions of the FFTW manual for R2HC/HC2R (halfcomplex order r0..r_{n/2},
// i_{ceil(n/2)-1}..i_1), REDFT00/01/10/11 and RODFT00/01/10/11.
//
// Fast path (n a power of two, kinds R2HC/HC2R/REDFT10/REDFT01/RODFT10/RODFT01): a tile of TP
// pencils is staged in shared memory; a real transform of length n is done as one complex
// radix-2 transform of length n/2 in place (DIF forward: natural -> bit-reversed; DIT backward:
// bit-reversed -> natural, so no reordering pass is needed), with the even/odd untangling step and
// the DCT/DST pre/post twiddles fused into the global store (forward) or global load (backward).
// Every pencil is read once and written once: 16 B per point.
// Generic path (any n, any kind): dense matrix-vector product against the precomputed n x n
// transform matrix from shared memory; used for short or non-power-of-two axes.
#include <algorithm>
#include <cmath>

#include "common.h"
#include "fft_radix.cuh"
#include "fft_row.cuh"

namespace snac {

// ------------------------------------------------------------------------------------------
// host: kind selection, normalisation, eigenvalues, dense matrices
// ------------------------------------------------------------------------------------------
void find_fft(const char bc[2], bool centered, int &kf, int &kb, double norm[2]) {
  // src/fft.f90:96-149
  const std::string b{bc[0], bc[1]};
  norm[0] = 2.0; norm[1] = 0.0;
  if (b == "FF") { kf = K_R2HC; kb = K_HC2R; norm[0] = 1.0; norm[1] = 0.0; return; }
  if (centered) {
    if (b == "NN") { kf = K_REDFT10; kb = K_REDFT01; }
    else if (b == "DD") { kf = K_RODFT10; kb = K_RODFT01; }
    else if (b == "ND") { kf = K_REDFT11; kb = K_REDFT11; }
    else if (b == "DN") { kf = K_RODFT11; kb = K_RODFT11; }
    else throw Error("find_fft: unsupported boundary-condition pair " + b);
  } else {
    if (b == "NN") { kf = K_REDFT00; kb = K_REDFT00; norm[1] = -1.0; }
    else if (b == "DD") { kf = K_RODFT00; kb = K_RODFT00; norm[1] = 1.0; }
    else if (b == "ND") { kf = K_REDFT10; kb = K_REDFT01; }
    else if (b == "DN") { kf = K_RODFT01; kb = K_RODFT10; }
    else throw Error("find_fft: unsupported boundary-condition pair " + b);
  }
}

std::vector<double> r2r_eigenvalues(int n, const char bc[2], bool centered) {
  // src/fft.f90:151-189.  For 'DN' the reference leaves lambda undefined; the 'ND' values are the
  // mathematically correct ones and are used here.
  const std::string b{bc[0], bc[1]};
  const double pi = std::acos(-1.0);
  std::vector<double> lam(n, 0.0);
  for (int l = 1; l <= n; ++l) {
    double s = 0.0;
    if (b == "FF") s = std::sin((l - 1.0) * pi / (1.0 * n));
    else if (b == "NN") s = std::sin((l - 1.0) * pi / (2.0 * n));
    else if (b == "DD") {
      if (centered) s = std::sin(l * pi / (2.0 * n));
      else s = (l <= n - 1) ? std::sin(l * pi / (2.0 * n)) : 0.0;
    } else if (b == "ND" || b == "DN") s = std::sin((2.0 * l - 1.0) * pi / (4.0 * n));
    else throw Error("eigenvalues: unsupported boundary-condition pair " + b);
    lam[l - 1] = -4.0 * s * s;
  }
  return lam;
}

static double cospi_frac(long a, long b) {  // cos(pi*a/b), argument reduced exactly
  a %= 2 * b;
  if (a < 0) a += 2 * b;
  return std::cos(std::acos(-1.0L) * (long double)a / (long double)b);
}
static double sinpi_frac(long a, long b) {
  a %= 2 * b;
  if (a < 0) a += 2 * b;
  return std::sin(std::acos(-1.0L) * (long double)a / (long double)b);
}

static std::vector<double> dense_matrix(int kind, int N) {
  std::vector<double> M((size_t)N * N, 0.0);
  auto at = [&](int k, int j) -> double & { return M[(size_t)k * N + j]; };
  for (int k = 0; k < N; ++k)
    for (int j = 0; j < N; ++j) {
      double v = 0.0;
      switch (kind) {
        case K_R2HC:
          if (k <= N / 2) v = cospi_frac(2L * j * k, N);
          else v = -sinpi_frac(2L * j * (N - k), N);
          break;
        case K_HC2R: {  // x_k = sum_j hc_j * (...)
          const int jj = j;  // halfcomplex index
          if (jj == 0) v = 1.0;
          else if (2 * jj == N) v = cospi_frac((long)k, 1);
          else if (jj <= N / 2) v = 2.0 * cospi_frac(2L * k * jj, N);
          else v = -2.0 * sinpi_frac(2L * k * (N - jj), N);
        } break;
        case K_REDFT00:
          if (N == 1) v = 1.0;
          else if (j == 0) v = 1.0;
          else if (j == N - 1) v = (k % 2) ? -1.0 : 1.0;
          else v = 2.0 * cospi_frac((long)j * k, N - 1);
          break;
        case K_REDFT10: v = 2.0 * cospi_frac((2L * j + 1) * k, 2L * N); break;
        case K_REDFT01: v = (j == 0) ? 1.0 : 2.0 * cospi_frac((long)j * (2L * k + 1), 2L * N); break;
        case K_REDFT11: v = 2.0 * cospi_frac((2L * j + 1) * (2L * k + 1), 4L * N); break;
        case K_RODFT00: v = 2.0 * sinpi_frac((long)(j + 1) * (k + 1), N + 1); break;
        case K_RODFT10: v = 2.0 * sinpi_frac((2L * j + 1) * (k + 1), 2L * N); break;
        case K_RODFT01:
          v = (j == N - 1) ? ((k % 2) ? -1.0 : 1.0) : 2.0 * sinpi_frac((long)(j + 1) * (2L * k + 1), 2L * N);
          break;
        case K_RODFT11: v = 2.0 * sinpi_frac((2L * j + 1) * (2L * k + 1), 4L * N); break;
        default: throw Error("dense_matrix: unknown transform kind");
      }
      at(k, j) = v;
    }
  return M;
}

// twiddle tables of the fast kernels: m/2 complex e^{-2 pi i q/m} (m = n/2), then m/2+1 complex
// e^{-2 pi i k/n}, then m+1 complex e^{-i pi k/(2n)}
std::vector<double> r2r_fast_tables(int n) {
  const int m = n / 2;
  std::vector<double> t;
  for (int k = 0; k < m / 2; ++k) { t.push_back(cospi_frac(2L * k, m)); t.push_back(-sinpi_frac(2L * k, m)); }
  for (int k = 0; k <= m / 2; ++k) { t.push_back(cospi_frac(2L * k, n)); t.push_back(-sinpi_frac(2L * k, n)); }
  for (int k = 0; k <= m; ++k) { t.push_back(cospi_frac(k, 2L * n)); t.push_back(-sinpi_frac(k, 2L * n)); }
  return t;
}

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

void r2r_plan_init(R2RPlan &pl, int axis, int n, const char bc[2], bool centered) {
  pl.axis = axis;
  pl.n = n;
  const int q = (!(bc[0] == 'F' && bc[1] == 'F') && !centered) ? 1 : 0;  // src/fft.f90:31-32
  pl.nt = n - q;
  double norm[2];
  find_fft(bc, centered, pl.kind_fwd, pl.kind_bwd, norm);
  pl.normfft = 1.0 / (norm[0] * (n + norm[1] - q));
  auto fast_kind = [](int k) {
    return k == K_R2HC || k == K_HC2R || k == K_REDFT10 || k == K_REDFT01 || k == K_RODFT10 || k == K_RODFT01;
  };
  pl.fast = (q == 0) && is_pow2(n) && n >= 8 && n <= 4096 && fast_kind(pl.kind_fwd) && fast_kind(pl.kind_bwd);
  if (pl.fast) {
    const std::vector<double> t = r2r_fast_tables(n);
    pl.twiddle.upload(t);
  } else {
    if ((size_t)pl.nt * pl.nt > (size_t)1 << 24) throw Error("transform length too large for the generic path");
    pl.mat_fwd.upload(dense_matrix(pl.kind_fwd, pl.nt));
    pl.mat_bwd.upload(dense_matrix(pl.kind_bwd, pl.nt));
  }
}

// ------------------------------------------------------------------------------------------
// device
// ------------------------------------------------------------------------------------------
// pencil p of tile `tile` -> offsets of its first element in the two arrays; false if out of range
__device__ inline bool pencil_base(const R2RArgs &a, int tile, int p, long &ib, long &ob) {
  const int ntp = (a.n[a.pd] + a.tp - 1) / a.tp;
  const int tpi = tile % ntp, o = tile / ntp;
  const int ip = tpi * a.tp + p;
  if (ip >= a.n[a.pd]) return false;
  ib = ip * a.ist[a.pd] + o * a.ist[a.od];
  ob = ip * a.ost[a.pd] + o * a.ost[a.od];
  return true;
}

// iterate work items (p, t) with the lane-fastest index chosen for coalescing

#define FOR_ITEMS(LANE_T_, NT_, ...)                                                     \
  {                                                                                      \
    const int tot_ = (NT_) * tp;                                                         \
    for (int it_ = threadIdx.x; it_ < tot_; it_ += blockDim.x) {                         \
      int p, t;                                                                          \
      if (LANE_T_) { t = it_ % (NT_); p = it_ / (NT_); } else { p = it_ % tp; t = it_ / tp; } \
      __VA_ARGS__                                                                        \
    }                                                                                    \
  }

__global__ void __launch_bounds__(256) r2r_dense_kernel(const R2RArgs a) {
  extern __shared__ double T[];
  const int tp = a.tp, tpp = a.tp + 1, nt = a.nt;
  const long isa = a.ist[a.axis], osa = a.ost[a.axis];
  const int nfull = a.n[a.axis];
  FOR_ITEMS(a.in_lane_t, nfull, {
    long ib, ob;
    if (pencil_base(a, blockIdx.x, p, ib, ob)) T[t * tpp + p] = R2R_IN_AT(t, ib, isa);
  })
  __syncthreads();
  FOR_ITEMS(a.out_lane_t, nfull, {
    long ib, ob;
    if (pencil_base(a, blockIdx.x, p, ib, ob)) {
      double acc;
      if (t < nt) {
        acc = 0.0;
        const double *row = a.tab + (size_t)t * nt;
        for (int j = 0; j < nt; ++j) acc += __ldg(row + j) * T[j * tpp + p];
      } else {
        acc = T[t * tpp + p];  // entries beyond the transform length (staggered, q = 1) pass through
      }
      R2R_OUT_AT(t, ob, osa) = acc * a.scale;
    }
  })
  if (a.out_split) __threadfence_system();   // stores into peers' slabs are performed before the kernel completes (bound.cu)
}

__device__ inline int brev(int k, int bits) { return (int)(__brev((unsigned)k) >> (32 - bits)); }

// complex helpers on (re, im) pairs
struct C2 { double x, y; };
__device__ inline C2 cmul(C2 a, C2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

template <bool FWD>
__global__ void __launch_bounds__(256) r2r_fast_kernel(const R2RArgs a) {
  extern __shared__ double T[];
  const int tp = a.tp, tpp = a.tp + 1, n = a.nt, m = n >> 1;
  const int bits = 31 - __clz(m);
  const long isa = a.ist[a.axis], osa = a.ost[a.axis];
  const double *tw = a.tab;                 // e^{-2 pi i q/m}, q < m/2
  const double *wn = a.tab + 2 * (m / 2);   // e^{-2 pi i k/n}, k <= m/2
  const double *wd = wn + 2 * (m / 2 + 1);  // e^{-i pi k/(2n)}, k <= m
  const bool makhoul = (a.kind == K_REDFT10 || a.kind == K_RODFT10 || a.kind == K_REDFT01 || a.kind == K_RODFT01);
  const bool sine = (a.kind == K_RODFT10 || a.kind == K_RODFT01);
  const int nk = m / 2 + 1;

  if (FWD) {
    // 1. load (with the even/odd permutation of the DCT-II algorithm and the (-1)^j of DST-II)
    FOR_ITEMS(a.in_lane_t, n, {
      long ib, ob;
      if (pencil_base(a, blockIdx.x, p, ib, ob)) {
        double v = R2R_IN_AT(t, ib, isa);
        int slot = t;
        if (makhoul) {
          if (sine && (t & 1)) v = -v;
          slot = (t & 1) ? (n - 1 - (t >> 1)) : (t >> 1);
        }
        T[slot * tpp + p] = v;
      }
    })
    __syncthreads();
    // 2. DIF radix-2 stages on z_j = T[2j] + i T[2j+1]
    for (int s = m >> 1; s >= 1; s >>= 1) {
      const int tot = (m >> 1) * tp;
      const int tws = (m >> 1) / s;  // twiddle stride: W_{2s}^t = tw[t*m/(2s)]
      for (int it = threadIdx.x; it < tot; it += blockDim.x) {
        const int p = it % tp, bf = it / tp;
        const int t = bf & (s - 1), g = (bf - t) << 1;
        const int i0 = g + t, i1 = i0 + s;
        double *r0 = T + (2 * i0) * tpp + p, *r1 = T + (2 * i1) * tpp + p;
        const C2 x0{r0[0], r0[tpp]}, x1{r1[0], r1[tpp]};
        const C2 d{x0.x - x1.x, x0.y - x1.y};
        const C2 w{__ldg(tw + 2 * t * tws), __ldg(tw + 2 * t * tws + 1)};
        const C2 y1 = cmul(d, w);
        r0[0] = x0.x + x1.x; r0[tpp] = x0.y + x1.y;
        r1[0] = y1.x; r1[tpp] = y1.y;
      }
      __syncthreads();
    }
    // 3. untangle + post-twiddle, written straight to global memory
    FOR_ITEMS(a.out_lane_t, nk, {
      long ib, ob;
      const int k = t;
      if (pencil_base(a, blockIdx.x, p, ib, ob)) {
        const int bk = brev(k, bits), bmk = brev((m - k) & (m - 1), bits);
        const C2 zk{T[(2 * bk) * tpp + p], T[(2 * bk + 1) * tpp + p]};
        const C2 zm{T[(2 * bmk) * tpp + p], T[(2 * bmk + 1) * tpp + p]};
        const C2 e{0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y)};
        const C2 o{0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x)};   // -i/2 (zk - conj zm)
        const C2 w{__ldg(wn + 2 * k), __ldg(wn + 2 * k + 1)};
        const C2 wo = cmul(w, o);
        const C2 xk{e.x + wo.x, e.y + wo.y};        // X[k]
        const C2 xmk{e.x - wo.x, -(e.y - wo.y)};    // X[m-k]
        const int ka = k, kb = m - k;               // the two output wavenumbers (k=0: 0 and m)
        const double sc = a.scale;
#define OUT_AT(idx) R2R_OUT_AT(idx, ob, osa)
        if (!makhoul) {
          OUT_AT(ka) = sc * xk.x;
          if (k > 0) OUT_AT(n - ka) = sc * xk.y;
          OUT_AT(kb) = sc * xmk.x;
          if (k > 0 && kb < m) OUT_AT(n - kb) = sc * xmk.y;
        } else {
          // X_c[k] = 2 Re(e^{-i pi k/2n} V[k]), X_c[n-k] = -2 Im(e^{-i pi k/2n} V[k])
          {
            const double c = __ldg(wd + 2 * ka), s_ = -__ldg(wd + 2 * ka + 1);
            const double v0 = 2.0 * (c * xk.x + s_ * xk.y), v1 = 2.0 * (s_ * xk.x - c * xk.y);
            const int o0 = sine ? (n - 1 - ka) : ka;
            OUT_AT(o0) = sc * v0;
            if (ka > 0) { const int o1 = sine ? (ka - 1) : (n - ka); OUT_AT(o1) = sc * v1; }
          }
          {
            const double c = __ldg(wd + 2 * kb), s_ = -__ldg(wd + 2 * kb + 1);
            const double v0 = 2.0 * (c * xmk.x + s_ * xmk.y), v1 = 2.0 * (s_ * xmk.x - c * xmk.y);
            const int o0 = sine ? (n - 1 - kb) : kb;
            OUT_AT(o0) = sc * v0;
            if (kb < m) { const int o1 = sine ? (kb - 1) : (n - kb); OUT_AT(o1) = sc * v1; }
          }
        }
      }
    })
  } else {
    // 1. global load + pre-twiddle + inverse untangle into bit-reversed slots
    FOR_ITEMS(a.in_lane_t, nk, {
      long ib, ob;
      const int k = t;
      if (pencil_base(a, blockIdx.x, p, ib, ob)) {
#define IN_AT(idx) R2R_IN_AT(idx, ib, isa)
        const int kb = m - k;
        C2 xk, xm;  // X[k], X[m-k]
        if (!makhoul) {
          xk.x = IN_AT(k);  xk.y = (k > 0) ? IN_AT(n - k) : 0.0;
          xm.x = IN_AT(kb); xm.y = (kb < m) ? IN_AT(n - kb) : 0.0;
        } else {
          auto getx = [&](int kk) -> C2 {
            const double c = __ldg(wd + 2 * kk), s_ = -__ldg(wd + 2 * kk + 1);
            const int i0 = sine ? (n - 1 - kk) : kk;
            const double a0 = IN_AT(i0);
            double a1 = 0.0;
            if (kk > 0) { const int i1 = sine ? (kk - 1) : (n - kk); a1 = IN_AT(i1); }
            return C2{c * a0 + s_ * a1, s_ * a0 - c * a1};
          };
          xk = getx(k);
          xm = getx(kb);
        }
        // Z[k] = (X[k] + conj X[m-k]) + i e^{+2 pi i k/n} (X[k] - conj X[m-k])
        const C2 w{__ldg(wn + 2 * k), -__ldg(wn + 2 * k + 1)};
        const C2 sum{xk.x + xm.x, xk.y - xm.y}, dif{xk.x - xm.x, xk.y + xm.y};
        const C2 wd_ = cmul(w, dif);
        const C2 zk{sum.x - wd_.y, sum.y + wd_.x};
        const int bk = brev(k & (m - 1), bits);
        T[(2 * bk) * tpp + p] = zk.x; T[(2 * bk + 1) * tpp + p] = zk.y;
        if (k > 0 && k != kb) {
          // Z[m-k] = conj( (X[k] + conj X[m-k]) - i e^{+2 pi i k/n}(X[k] - conj X[m-k]) )
          const C2 zm{sum.x + wd_.y, -(sum.y - wd_.x)};
          const int bm = brev(kb, bits);
          T[(2 * bm) * tpp + p] = zm.x; T[(2 * bm + 1) * tpp + p] = zm.y;
        }
      }
    })
    __syncthreads();
    // 2. DIT radix-2 stages (inverse sign)
    for (int s = 1; s < m; s <<= 1) {
      const int tot = (m >> 1) * tp;
      const int tws = (m >> 1) / s;
      for (int it = threadIdx.x; it < tot; it += blockDim.x) {
        const int p = it % tp, bf = it / tp;
        const int t = bf & (s - 1), g = (bf - t) << 1;
        const int i0 = g + t, i1 = i0 + s;
        double *r0 = T + (2 * i0) * tpp + p, *r1 = T + (2 * i1) * tpp + p;
        const C2 x0{r0[0], r0[tpp]}, x1{r1[0], r1[tpp]};
        const C2 w{__ldg(tw + 2 * t * tws), -__ldg(tw + 2 * t * tws + 1)};
        const C2 y1 = cmul(x1, w);
        r0[0] = x0.x + y1.x; r0[tpp] = x0.y + y1.y;
        r1[0] = x0.x - y1.x; r1[tpp] = x0.y - y1.y;
      }
      __syncthreads();
    }
    // 3. store (inverse of the forward load mapping)
    FOR_ITEMS(a.out_lane_t, n, {
      long ib, ob;
      if (pencil_base(a, blockIdx.x, p, ib, ob)) {
        int slot = t;
        if (makhoul) slot = (t & 1) ? (n - 1 - (t >> 1)) : (t >> 1);
        double v = T[slot * tpp + p] * a.scale;
        if (makhoul && sine && (t & 1)) v = -v;
        R2R_OUT_AT(t, ob, osa) = v;
      }
    })
  }
  if (a.out_split) __threadfence_system();
}

// register-radix kernels (fft_radix.cuh): one CTA per tile of 32 pencils
template <int R1, int R2, bool FWD, bool SPLIT, bool STAGE = false>
__global__ void __launch_bounds__(radix::Plan<R1, R2>::THREADS, FWD ? radix::Plan<R1, R2>::MINB_FWD : 0) r2r_radix_kernel(const R2RArgs a) {
  extern __shared__ __align__(16) double2 Zs[];
  const radix::Thread T = radix::thread_of<radix::Plan<R1, R2>::TP>(a, threadIdx.x, blockDim.x, blockIdx.x);
  if (FWD) {
    if (STAGE) {
      radix::stage_io<R1, R2>(a, Zs, threadIdx.x, blockDim.x, blockIdx.x, true);
      __syncthreads();
    }
    radix::fwd_pass1<R1, R2, SPLIT, STAGE>(a, Zs, T);
    __syncthreads();
    radix::run_pass<R1, R2, false>(a, Zs, T);
    __syncthreads();
    radix::fwd_untangle<R1, R2, SPLIT>(a, Zs, T);
    // the forward transform stores into the owners' slabs, possibly over NVLink: performed there before the kernel
    // completes.  One system-scope fence per CTA after the barrier (fences are cumulative: the stores of every thread
    // that arrived at the barrier are ordered before what thread 0 does next) instead of one per thread: 256 threads
    // waiting for the fabric's acknowledgements kept the CTA's registers and shared memory busy
    if (SPLIT) { __syncthreads(); if (threadIdx.x == 0) __threadfence_system(); }
  } else {
    radix::bwd_tangle<R1, R2, SPLIT>(a, Zs, T);
    __syncthreads();
    radix::run_pass<R1, R2, true>(a, Zs, T);
    __syncthreads();
    radix::bwd_pass2<R1, R2, SPLIT, STAGE>(a, Zs, T);
    if (STAGE) {
      __syncthreads();
      radix::stage_io<R1, R2>(a, Zs, threadIdx.x, blockDim.x, blockIdx.x, false);
    }
  }
}

template <int R1, int R2, bool FWD, bool SPLIT, bool STAGE>
static void radix_go1(cudaStream_t st, const R2RArgs &a, long ntiles) {
  using P = radix::Plan<R1, R2>;
  static bool attr = false;
  if (!attr && P::SMEM > 40 * 1024) {
    SNAC_CUDA(cudaFuncSetAttribute(r2r_radix_kernel<R1, R2, FWD, SPLIT, STAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM));
    attr = true;
  }
  r2r_radix_kernel<R1, R2, FWD, SPLIT, STAGE><<<(unsigned)ntiles, P::THREADS, P::SMEM, st>>>(a);
}

template <int R1, int R2>
static void radix_go(cudaStream_t st, const R2RArgs &a, int dir, long ntiles, bool stage) {
  const bool split = a.in_split != 0 || a.out_split != 0;
  if (stage) {
    if (dir > 0) { if (split) radix_go1<R1, R2, true, true, true>(st, a, ntiles); else radix_go1<R1, R2, true, false, true>(st, a, ntiles); }
    else         { if (split) radix_go1<R1, R2, false, true, true>(st, a, ntiles); else radix_go1<R1, R2, false, false, true>(st, a, ntiles); }
  } else {
    if (dir > 0) { if (split) radix_go1<R1, R2, true, true, false>(st, a, ntiles); else radix_go1<R1, R2, true, false, false>(st, a, ntiles); }
    else         { if (split) radix_go1<R1, R2, false, true, false>(st, a, ntiles); else radix_go1<R1, R2, false, false, false>(st, a, ntiles); }
  }
}

// transform lengths served by the radix kernels: n = 2 R1 R2
static bool radix_launch(cudaStream_t st, const R2RArgs &a, int dir, long ntiles, bool stage) {
  switch (a.nt) {
    case 32: radix_go<4, 4>(st, a, dir, ntiles, stage); return true;
    case 64: radix_go<8, 4>(st, a, dir, ntiles, stage); return true;
    case 128: radix_go<8, 8>(st, a, dir, ntiles, stage); return true;
    case 256: radix_go<16, 8>(st, a, dir, ntiles, stage); return true;
    case 512: radix_go<16, 16>(st, a, dir, ntiles, stage); return true;
    default: return false;
  }
}

void r2r_execute(Ctx &ctx, const R2RPlan &pl, int dir, const int n3[3], const double *const *in, int in_split,
                 const long ist[3], double *const *out, int out_split, const long ost[3], int chunk, double scale,
                 bool cyclic) {
  R2RArgs a;
  for (int p = 0; p < 16; ++p) { a.inp[p] = nullptr; a.outp[p] = nullptr; }
  for (int p = 0; p < std::max(in_split, 1); ++p) a.inp[p] = in[p];
  for (int p = 0; p < std::max(out_split, 1); ++p) a.outp[p] = out[p];
  const int mode = cyclic ? 2 : 1;
  a.in_split = in_split > 1 ? mode : 0; a.out_split = out_split > 1 ? mode : 0; a.chunk = std::max(chunk, 1);
  a.ns_mask = 0; a.ns_shift = 0;
  if (cyclic) {
    const int ns = std::max(in_split, out_split);
    if (ns < 1 || (ns & (ns - 1))) throw Error("r2r_execute: a cyclic split needs a power-of-two number of pieces");
    a.ns_mask = ns - 1;
    while ((1 << a.ns_shift) < ns) ++a.ns_shift;
  }
  for (int d = 0; d < 3; ++d) { a.n[d] = n3[d]; a.ist[d] = ist[d]; a.ost[d] = ost[d]; }
  a.axis = pl.axis;
  a.nt = pl.nt;
  a.kind = dir > 0 ? pl.kind_fwd : pl.kind_bwd;
  a.scale = scale;
  const int n = n3[pl.axis];
  if (n != pl.n) throw Error("r2r_execute: box size differs from the planned transform length");
  a.pd = pl.axis == 0 ? 1 : 0;
  a.od = 3 - pl.axis - a.pd;
  a.in_lane_t = (ist[pl.axis] == 1);
  a.out_lane_t = (ost[pl.axis] == 1);
  int tp = 16;
  while (tp > 1 && (size_t)n * (tp + 1) * sizeof(double) > 200 * 1024) tp >>= 1;
  while (tp > 1 && n3[a.pd] < tp) tp >>= 1;
  a.tp = tp;
  const long ntiles = (long)((n3[a.pd] + tp - 1) / tp) * n3[a.od];
  const size_t smem = (size_t)n * (tp + 1) * sizeof(double);
  static const bool staged_only = getenv("SNAC_FFT_STAGED") && atoi(getenv("SNAC_FFT_STAGED")) != 0;   // test hook
  if (pl.fast && a.in_lane_t && a.out_lane_t && !staged_only) {   // the axis is contiguous on both sides: one warp per row
    a.tab = pl.twiddle.p();
    if (rowfft::row_launch(ctx.stream, a, dir)) {
      ctx.launches++;
      SNAC_CUDA(cudaGetLastError());
      return;
    }
  }
  // the radix kernels: lanes along the pencils on both sides, or -- the transform axis at unit stride on the physical
  // side (input of a forward, output of a backward transform), the pencils at unit stride on the spectral side -- with
  // the physical tile staged through shared memory
  const bool phys_lane_t = dir > 0 ? (a.in_lane_t && !a.out_lane_t && in_split <= 1 && ost[a.pd] == 1)
                                   : (a.out_lane_t && !a.in_lane_t && out_split <= 1 && ist[a.pd] == 1);
  if (pl.fast && ((!a.in_lane_t && !a.out_lane_t) || phys_lane_t) && !staged_only) {
    a.tab = pl.twiddle.p();
    a.tp = radix::tp_of(a.nt);
    const long nt16 = (long)((n3[a.pd] + a.tp - 1) / a.tp) * n3[a.od];
    if (radix_launch(ctx.stream, a, dir, nt16, phys_lane_t)) {
      ctx.launches++;
      SNAC_CUDA(cudaGetLastError());
      return;
    }
    a.tp = tp;
  }
  if (pl.fast) {
    a.tab = pl.twiddle.p();
    if (dir > 0) {
      if (smem > 40 * 1024) SNAC_CUDA(cudaFuncSetAttribute(r2r_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      r2r_fast_kernel<true><<<(unsigned)ntiles, 256, smem, ctx.stream>>>(a);
    } else {
      if (smem > 40 * 1024) SNAC_CUDA(cudaFuncSetAttribute(r2r_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      r2r_fast_kernel<false><<<(unsigned)ntiles, 256, smem, ctx.stream>>>(a);
    }
  } else {
    a.tab = dir > 0 ? pl.mat_fwd.p() : pl.mat_bwd.p();
    if (smem > 40 * 1024) SNAC_CUDA(cudaFuncSetAttribute(r2r_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    r2r_dense_kernel<<<(unsigned)ntiles, 256, smem, ctx.stream>>>(a);
  }
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

}  // namespace snac
