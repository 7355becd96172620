// K3b/K5, register-radix form of the fast real-to-real transforms (included by fft.cu).
//
// Same mathematics as r2r_fast_kernel (a real transform of length n as one complex transform of
// length m = n/2 plus an untangling step; DCT/DST-II/III through Makhoul's even/odd permutation and
// a post/pre-twiddle), different data movement: the complex transform is done in two in-place
// mixed-radix passes, m = R1 * R2 with radices up to 16 held in registers,
//
//   forward :  global -> [radix-R1 butterflies at stride m/R1, twiddle] -> smem
//              smem   -> [radix-R2 butterflies on contiguous runs]      -> smem   (digit-reversed)
//              smem   -> untangle + post-twiddle                        -> global
//   backward:  global -> pre-twiddle + inverse untangle                 -> smem   (digit-reversed)
//              smem   -> [radix-R2 on contiguous runs, twiddle]         -> smem
//              smem   -> [radix-R1 at stride m/R1]                      -> global
//
// so that (i) every thread has all the global loads of its butterfly (2 R1 rows) in flight at once --
// the staged kernel issued them one by one in front of a shared-memory store and was bound by that
// latency -- and (ii) an element crosses shared memory twice instead of eight times.  Consecutive
// lanes hold consecutive pencils (the unit-stride direction of both arrays), a complex element is
// one 16-byte shared-memory word.  Used when neither side has the transform axis at unit stride
// (FFT_AXIS = 2, 3); otherwise the staged kernel does the transposition through shared memory.
//
// The phase functions are __host__ __device__ so that tests/host/fft_emul.cu can run them thread by
// thread on the CPU against the dense transform definitions (no GPU needed for the index algebra).
#pragma once
#include "common.h"

namespace snac {

struct R2RArgs {
  // first interior element of the input / output array; when the wavenumber axis of that side is
  // split over processes (nsplit > 1), one base per owning process, chosen by index / chunk
  const double *inp[16];
  double *outp[16];
  int in_split, out_split, chunk;   // see R2R_WHO / R2R_POS
  int ns_mask, ns_shift;
  long ist[3], ost[3];  // strides of the physical directions x,y,z in the input / output arrays
  int n[3];
  int axis;
  int nt;            // transform length
  int tp;            // pencils per tile
  int pd, od;        // direction along which a tile gathers pencils; the remaining direction
  int kind;
  int in_lane_t, out_lane_t;  // consecutive lanes walk along the transform axis (it has stride 1)
  double scale;
  const double *tab; // twiddles (fast) or dense matrix (generic)
};

// element t of a pencil whose first element is at offset `base` and whose axis stride is `st`.
// split = 0: one array.  split = 1: the axis is cut into consecutive chunks of `chunk` indices, one
// array per chunk (bases shifted by the chunk origin).  split = 2: cyclic, index t lives in array
// t mod ns at position t div ns (ns a power of two): every process owns low and high wavenumbers
// alike, which balances the solver's work per process (low wavenumbers need the most cycles).
#define R2R_WHO(split, t) ((split) == 0 ? 0 : ((split) == 1 ? (t) / a.chunk : ((t) & a.ns_mask)))
#define R2R_POS(split, t) ((split) == 2 ? ((t) >> a.ns_shift) : (t))
#define R2R_IN_AT(t, base, st) a.inp[R2R_WHO(a.in_split, t)][(base) + (long)R2R_POS(a.in_split, t) * (st)]
#define R2R_OUT_AT(t, base, st) a.outp[R2R_WHO(a.out_split, t)][(base) + (long)R2R_POS(a.out_split, t) * (st)]
// the radix kernels are compiled twice: SPLIT = false (one process: one array on either side) keeps the
// address arithmetic of the unrolled butterflies free of the selects above
#define R2R_IN_ATS(t, base, st) (SPLIT ? R2R_IN_AT(t, base, st) : a.inp[0][(base) + (t) * (st)])
#define R2R_OUT_ATS(t, base, st) (SPLIT ? R2R_OUT_AT(t, base, st) : a.outp[0][(base) + (t) * (st)])

#if defined(__CUDA_ARCH__)
#define SNAC_LDG(p) __ldg(p)
#else
#define SNAC_LDG(p) (*(p))
#endif
#define SNAC_HD __host__ __device__ __forceinline__

namespace radix {

struct Cx { double x, y; };
SNAC_HD Cx operator+(Cx a, Cx b) { return {a.x + b.x, a.y + b.y}; }
SNAC_HD Cx operator-(Cx a, Cx b) { return {a.x - b.x, a.y - b.y}; }
SNAC_HD Cx cmul(Cx a, Cx b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

// cos(2 pi k / 16), sin(2 pi k / 16), k = 0..15
#define SNAC_C16 {1.0, 0.92387953251128673848, 0.70710678118654752440, 0.38268343236508977173, 0.0, -0.38268343236508977173, \
                  -0.70710678118654752440, -0.92387953251128673848, -1.0, -0.92387953251128673848, -0.70710678118654752440, \
                  -0.38268343236508977173, 0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848}
#define SNAC_S16 {0.0, 0.38268343236508977173, 0.70710678118654752440, 0.92387953251128673848, 1.0, 0.92387953251128673848, \
                  0.70710678118654752440, 0.38268343236508977173, 0.0, -0.38268343236508977173, -0.70710678118654752440, \
                  -0.92387953251128673848, -1.0, -0.92387953251128673848, -0.70710678118654752440, -0.38268343236508977173}

// v *= exp(-/+ 2 pi i K / 16) (INV: +), K a compile-time constant after unrolling
template <bool INV>
SNAC_HD Cx rot16(Cx v, int K) {
  constexpr double c16[16] = SNAC_C16, s16[16] = SNAC_S16;
  K &= 15;
  if (K == 0) return v;
  if (K == 4) return INV ? Cx{-v.y, v.x} : Cx{v.y, -v.x};
  if (K == 8) return Cx{-v.x, -v.y};
  if (K == 12) return INV ? Cx{v.y, -v.x} : Cx{-v.y, v.x};
  const double c = c16[K], s = INV ? s16[K] : -s16[K];
  return Cx{v.x * c - v.y * s, v.x * s + v.y * c};
}

__host__ __device__ constexpr int brev_c(int k, int bits) {
  int r = 0;
  for (int b = 0; b < bits; ++b) r |= ((k >> b) & 1) << (bits - 1 - b);
  return r;
}
__host__ __device__ constexpr int log2_c(int n) { return n <= 1 ? 0 : 1 + log2_c(n / 2); }

// unnormalised DFT of length R (a power of two <= 16) in registers, natural order in and out:
//   v[k] <- sum_j v[j] exp(-/+ 2 pi i j k / R)
template <int R, bool INV>
SNAC_HD void dft(Cx (&v)[R]) {
#pragma unroll
  for (int len = R; len >= 2; len >>= 1) {   // radix-2 DIF stages: natural -> bit-reversed
    const int half = len >> 1;
#pragma unroll
    for (int b = 0; b < R; b += len) {
#pragma unroll
      for (int t = 0; t < half; ++t) {
        const Cx p = v[b + t], q = v[b + t + half];
        v[b + t] = p + q;
        v[b + t + half] = rot16<INV>(p - q, t * (16 / len));
      }
    }
  }
  constexpr int bits = log2_c(R);
#pragma unroll
  for (int k = 0; k < R; ++k) {   // undo the bit reversal (register renaming after unrolling)
    const int r = brev_c(k, bits);
    if (r > k) { const Cx t = v[k]; v[k] = v[r]; v[r] = t; }
  }
}

// pencils per tile = lanes along the unit-stride direction.  32 pencils (rows of 256 bytes) up to n = 256; at n = 512 a
// tile of 32 pencils fills 128 KB of shared memory: one CTA per SM, whose load, butterfly and store phases then run one
// after the other with nothing to overlap them.  16 pencils (64 KB, two to three CTAs per SM) measured on the channel
// (n = 512 along y, 1024 x 512 pencils): forward 4.28 -> 5.05 TB/s, backward 3.83 -> 4.88 TB/s.
#ifndef SNAC_FFT_TP_LONG
#define SNAC_FFT_TP_LONG 16
#endif
#ifndef SNAC_FFT_MINB_FROM
#define SNAC_FFT_MINB_FROM 256   // shortest transform whose forward kernel is held to SNAC_FFT_MINB_LONG CTAs per SM (64 KB tiles:
                                 // n = 256 with 32 pencils, n = 512 with 16; Taylor-Green 256^3: forward 3.8 -> 4.5 TB/s)
#endif
#ifndef SNAC_FFT_MINB_LONG
#define SNAC_FFT_MINB_LONG 3
#endif
__host__ __device__ constexpr int tp_of(int n) { return n >= 512 ? SNAC_FFT_TP_LONG : 32; }

// per-thread constants of a tile
struct Thread {
  int p, jt, tpp;      // pencil in the tile, index among the threads of that pencil, threads per pencil
  int ip, o;           // the pencil's indices along pd and od (0-based)
  bool valid;          // the pencil exists
  long ib, ob;         // offsets of the pencil's first element in the input / output arrays
};

template <int TP>
SNAC_HD Thread thread_of(const R2RArgs &a, int tid, int nthr, int tile) {
  Thread T;
  T.p = tid % TP; T.jt = tid / TP; T.tpp = nthr / TP;
  const int ntp = (a.n[a.pd] + TP - 1) / TP;
  const int tpi = tile % ntp, o = tile / ntp;
  const int ip = tpi * TP + T.p;
  T.valid = ip < a.n[a.pd];
  T.ip = ip; T.o = o;
  T.ib = ip * a.ist[a.pd] + o * a.ist[a.od];
  T.ob = ip * a.ost[a.pd] + o * a.ost[a.od];
  return T;
}

// e^{-2 pi i k / m} for 0 <= k < m from the table of the first half turn
SNAC_HD Cx tw_m(const double *tw, int k, int m) {
  const int h = m >> 1;
  const int kk = k >= h ? k - h : k;
  Cx w{SNAC_LDG(tw + 2 * kk), SNAC_LDG(tw + 2 * kk + 1)};
  if (k >= h) { w.x = -w.x; w.y = -w.y; }
  return w;
}

// Makhoul's permutation: slot s of the permuted sequence holds input element tof(s)
SNAC_HD int tof(int s, int n, bool makhoul) { return !makhoul ? s : (2 * s < n ? 2 * s : 2 * (n - 1 - s) + 1); }

template <int R1, int R2>
struct Plan {
  static constexpr int M = R1 * R2;       // complex length
  static constexpr int N = 2 * M;         // real length
  static constexpr int S1 = R2;           // stride of the first pass = length of the contiguous runs
  static constexpr int TP = tp_of(N);              // pencils per tile
  static constexpr int TPP = R1 < R2 ? R1 : R2;   // threads per pencil
  static constexpr int THREADS = TP * TPP;
  // CTAs per SM the register allocation of the forward kernel aims at (0: the compiler's choice).  n >= 256: three CTAs
  // of 64 KB; with several processes the stores of a tile cross NVLink and only other CTAs can cover that latency
  // (2 GPUs, channel: forward transform 4.97 -> 3.35 ms per step)
  static constexpr int MINB_FWD = N >= SNAC_FFT_MINB_FROM ? SNAC_FFT_MINB_LONG : 0;
  static constexpr size_t SMEM = (size_t)M * TP * sizeof(double2);
  SNAC_HD static int loc(int k) { return (k % R1) * S1 + k / R1; }   // where X[k] sits after the two passes
};

SNAC_HD bool is_makhoul(int kind) { return kind == K_REDFT10 || kind == K_RODFT10 || kind == K_REDFT01 || kind == K_RODFT01; }
SNAC_HD bool is_sine(int kind) { return kind == K_RODFT10 || kind == K_RODFT01; }

// ---- forward ---------------------------------------------------------------------------------
// STAGE: the tile was brought into Z by stage_in (the transform axis is at unit stride on the physical side): element e
// of the permuted sequence, sign included, sits where this thread's own output for the same e goes
template <int R1, int R2, bool SPLIT, bool STAGE = false>
SNAC_HD void fwd_pass1(const R2RArgs &a, double2 *Z, const Thread &T) {
  using P = Plan<R1, R2>;
  if (!T.valid) return;
  const bool makhoul = is_makhoul(a.kind), sine = is_sine(a.kind);
  const long isa = a.ist[a.axis];
  const double *tw = a.tab;
#pragma unroll
  for (int j0 = 0; j0 < P::S1; j0 += P::TPP) {
    const int j = j0 + T.jt;
    if (j >= P::S1) break;
    Cx v[R1];
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      const int e = j + r * P::S1;
      if (STAGE) {
        const double2 z = Z[e * P::TP + T.p];
        v[r] = Cx{z.x, z.y};
        continue;
      }
      const int t0 = tof(2 * e, P::N, makhoul), t1 = tof(2 * e + 1, P::N, makhoul);
      double x0 = R2R_IN_ATS(t0, T.ib, isa), x1 = R2R_IN_ATS(t1, T.ib, isa);
      if (sine && (t0 & 1)) x0 = -x0;
      if (sine && (t1 & 1)) x1 = -x1;
      v[r] = Cx{x0, x1};
    }
    dft<R1, false>(v);
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      const Cx y = q == 0 ? v[0] : cmul(v[q], tw_m(tw, j * q, P::M));
      Z[(j + q * P::S1) * P::TP + T.p] = make_double2(y.x, y.y);
    }
  }
}

template <int R1, int R2, bool INV>
SNAC_HD void run_pass(const R2RArgs &a, double2 *Z, const Thread &T) {   // radix R2 on the contiguous runs
  using P = Plan<R1, R2>;
  if (!T.valid) return;
  const double *tw = a.tab;
#pragma unroll
  for (int q0 = 0; q0 < R1; q0 += P::TPP) {
    const int q = q0 + T.jt;
    if (q >= R1) break;
    Cx v[R2];
#pragma unroll
    for (int r = 0; r < R2; ++r) {
      const double2 z = Z[(q * P::S1 + r) * P::TP + T.p];
      v[r] = Cx{z.x, z.y};
    }
    dft<R2, INV>(v);
#pragma unroll
    for (int s = 0; s < R2; ++s) {
      Cx y = v[s];
      if (INV && s > 0 && q > 0) {   // backward: twiddle between the passes, conjugated
        Cx w = tw_m(tw, s * q, P::M);
        w.y = -w.y;
        y = cmul(y, w);
      }
      Z[(q * P::S1 + s) * P::TP + T.p] = make_double2(y.x, y.y);
    }
  }
}

template <int R1, int R2, bool SPLIT>
SNAC_HD void fwd_untangle(const R2RArgs &a, const double2 *Z, const Thread &T) {
  using P = Plan<R1, R2>;
  if (!T.valid) return;
  constexpr int m = P::M, n = P::N, nk = m / 2 + 1;
  const bool makhoul = is_makhoul(a.kind), sine = is_sine(a.kind);
  const long osa = a.ost[a.axis], ob = T.ob;
  const double *wn = a.tab + 2 * (m / 2);   // e^{-2 pi i k/n}, k <= m/2
  const double *wd = wn + 2 * (m / 2 + 1);  // e^{-i pi k/(2n)}, k <= m
  const double sc = a.scale;
#pragma unroll
  for (int k0 = 0; k0 < nk; k0 += P::TPP) {
    const int k = k0 + T.jt;
    if (k >= nk) break;
    const double2 zk2 = Z[P::loc(k) * P::TP + T.p], zm2 = Z[P::loc((m - k) & (m - 1)) * P::TP + T.p];
    const Cx zk{zk2.x, zk2.y}, zm{zm2.x, zm2.y};
    const Cx e{0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y)};
    const Cx o{0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x)};   // -i/2 (zk - conj zm)
    const Cx w{SNAC_LDG(wn + 2 * k), SNAC_LDG(wn + 2 * k + 1)};
    const Cx wo = cmul(w, o);
    const Cx xk{e.x + wo.x, e.y + wo.y};        // X[k]
    const Cx xmk{e.x - wo.x, -(e.y - wo.y)};    // X[m-k]
    const int ka = k, kb = m - k;               // the two output wavenumbers (k=0: 0 and m)
#define SNAC_OUT_AT(idx) R2R_OUT_ATS(idx, ob, osa)
    if (!makhoul) {
      SNAC_OUT_AT(ka) = sc * xk.x;
      if (k > 0) SNAC_OUT_AT(n - ka) = sc * xk.y;
      SNAC_OUT_AT(kb) = sc * xmk.x;
      if (k > 0 && kb < m) SNAC_OUT_AT(n - kb) = sc * xmk.y;
    } else {
      // X_c[k] = 2 Re(e^{-i pi k/2n} V[k]), X_c[n-k] = -2 Im(e^{-i pi k/2n} V[k])
      {
        const double c = SNAC_LDG(wd + 2 * ka), s_ = -SNAC_LDG(wd + 2 * ka + 1);
        const double v0 = 2.0 * (c * xk.x + s_ * xk.y), v1 = 2.0 * (s_ * xk.x - c * xk.y);
        const int o0 = sine ? (n - 1 - ka) : ka;
        SNAC_OUT_AT(o0) = sc * v0;
        if (ka > 0) { const int o1 = sine ? (ka - 1) : (n - ka); SNAC_OUT_AT(o1) = sc * v1; }
      }
      {
        const double c = SNAC_LDG(wd + 2 * kb), s_ = -SNAC_LDG(wd + 2 * kb + 1);
        const double v0 = 2.0 * (c * xmk.x + s_ * xmk.y), v1 = 2.0 * (s_ * xmk.x - c * xmk.y);
        const int o0 = sine ? (n - 1 - kb) : kb;
        SNAC_OUT_AT(o0) = sc * v0;
        if (kb < m) { const int o1 = sine ? (kb - 1) : (n - kb); SNAC_OUT_AT(o1) = sc * v1; }
      }
    }
#undef SNAC_OUT_AT
  }
}

// ---- backward --------------------------------------------------------------------------------
// All the loads of a thread (four per wavenumber pair it untangles) are issued before the first of them is used:
// issued pair by pair in front of the shared-memory stores, a thread had four loads in flight and the n = 512
// transform (one CTA of 512 threads per SM) ran at 0.45 of the copy rate.
template <int R1, int R2, bool SPLIT>
SNAC_HD void bwd_tangle(const R2RArgs &a, double2 *Z, const Thread &T) {
  using P = Plan<R1, R2>;
  if (!T.valid) return;
  constexpr int m = P::M, n = P::N, nk = m / 2 + 1;
  constexpr int NIT = (nk + P::TPP - 1) / P::TPP;
  const bool makhoul = is_makhoul(a.kind), sine = is_sine(a.kind);
  const long isa = a.ist[a.axis], ib = T.ib;
  const double *wn = a.tab + 2 * (m / 2);
  const double *wd = wn + 2 * (m / 2 + 1);
#define SNAC_IN_AT(idx) R2R_IN_ATS(idx, ib, isa)
  double in[NIT][4];   // per pair: the two halfcomplex (or DCT/DST) entries of X[k], then those of X[m-k]
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int k = it * P::TPP + T.jt;
    in[it][0] = in[it][1] = in[it][2] = in[it][3] = 0.0;
    if (k >= nk) continue;
    const int kb = m - k;
    if (!makhoul) {
      in[it][0] = SNAC_IN_AT(k);
      if (k > 0) in[it][1] = SNAC_IN_AT(n - k);
      in[it][2] = SNAC_IN_AT(kb);
      if (kb < m) in[it][3] = SNAC_IN_AT(n - kb);
    } else {
      in[it][0] = SNAC_IN_AT(sine ? (n - 1 - k) : k);
      if (k > 0) in[it][1] = SNAC_IN_AT(sine ? (k - 1) : (n - k));
      in[it][2] = SNAC_IN_AT(sine ? (n - 1 - kb) : kb);
      if (kb > 0) in[it][3] = SNAC_IN_AT(sine ? (kb - 1) : (n - kb));
    }
  }
#undef SNAC_IN_AT
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int k = it * P::TPP + T.jt;
    if (k >= nk) continue;
    const int kb = m - k;
    Cx xk, xm;  // X[k], X[m-k]
    if (!makhoul) {
      xk = Cx{in[it][0], in[it][1]};
      xm = Cx{in[it][2], in[it][3]};
    } else {
      const double c0 = SNAC_LDG(wd + 2 * k), s0 = -SNAC_LDG(wd + 2 * k + 1);
      const double c1 = SNAC_LDG(wd + 2 * kb), s1 = -SNAC_LDG(wd + 2 * kb + 1);
      xk = Cx{c0 * in[it][0] + s0 * in[it][1], s0 * in[it][0] - c0 * in[it][1]};
      xm = Cx{c1 * in[it][2] + s1 * in[it][3], s1 * in[it][2] - c1 * in[it][3]};
    }
    // Z[k] = (X[k] + conj X[m-k]) + i e^{+2 pi i k/n} (X[k] - conj X[m-k])
    const Cx w{SNAC_LDG(wn + 2 * k), -SNAC_LDG(wn + 2 * k + 1)};
    const Cx sum{xk.x + xm.x, xk.y - xm.y}, dif{xk.x - xm.x, xk.y + xm.y};
    const Cx wd_ = cmul(w, dif);
    Z[P::loc(k & (m - 1)) * P::TP + T.p] = make_double2(sum.x - wd_.y, sum.y + wd_.x);
    if (k > 0 && k != kb)   // Z[m-k] = conj( (X[k] + conj X[m-k]) - i e^{+2 pi i k/n}(X[k] - conj X[m-k]) )
      Z[P::loc(kb) * P::TP + T.p] = make_double2(sum.x + wd_.y, -(sum.y - wd_.x));
  }
}

// STAGE: the results go back into the thread's own slots of Z (element e of the permuted sequence, scaled, sign
// included) and leave through stage_out
template <int R1, int R2, bool SPLIT, bool STAGE = false>
SNAC_HD void bwd_pass2(const R2RArgs &a, double2 *Z, const Thread &T) {   // radix R1 at stride S1, then the store
  using P = Plan<R1, R2>;
  if (!T.valid) return;
  const bool makhoul = is_makhoul(a.kind), sine = is_sine(a.kind);
  const long osa = a.ost[a.axis];
#pragma unroll
  for (int j0 = 0; j0 < P::S1; j0 += P::TPP) {
    const int j = j0 + T.jt;
    if (j >= P::S1) break;
    Cx v[R1];
#pragma unroll
    for (int q = 0; q < R1; ++q) {
      const double2 z = Z[(q * P::S1 + j) * P::TP + T.p];
      v[q] = Cx{z.x, z.y};
    }
    dft<R1, true>(v);
#pragma unroll
    for (int r = 0; r < R1; ++r) {
      const int e = j + r * P::S1;
      const int t0 = tof(2 * e, P::N, makhoul), t1 = tof(2 * e + 1, P::N, makhoul);
      double x0 = v[r].x * a.scale, x1 = v[r].y * a.scale;
      if (sine && (t0 & 1)) x0 = -x0;
      if (sine && (t1 & 1)) x1 = -x1;
      if (STAGE) { Z[e * P::TP + T.p] = make_double2(x0, x1); continue; }
      R2R_OUT_ATS(t0, T.ob, osa) = x0;
      R2R_OUT_ATS(t1, T.ob, osa) = x1;
    }
  }
}

// ---- transform axis at unit stride on the physical side (FFT_AXIS = 1) -----------------------------------------
// The pencils of a tile are rows of the physical array: lanes along the pencils would read 8 bytes per row.  The
// tile goes through shared memory instead: a quarter warp takes 4 consecutive elements (64 contiguous bytes) of 8
// pencils and writes them as one conflict-free 128-byte shared-memory wavefront per element, into the slots from which
// (into which) the first (last) radix pass of the owning thread reads (writes) them.
template <int R1, int R2>
SNAC_HD void stage_io(const R2RArgs &a, double2 *Z, int tid, int nthr, int tile, bool in) {
  using P = Plan<R1, R2>;
  constexpr int TP = P::TP, M = P::M, N = P::N;
  const bool makhoul = is_makhoul(a.kind), sine = is_sine(a.kind);
  const int ntp = (a.n[a.pd] + TP - 1) / TP;
  const int ip0 = (tile % ntp) * TP, o = tile / ntp;
  const long st_ax = in ? a.ist[a.axis] : a.ost[a.axis];
  const long st_pd = in ? a.ist[a.pd] : a.ost[a.pd], st_od = in ? a.ist[a.od] : a.ost[a.od];
  for (int idx = tid; idx < M * TP; idx += nthr) {
    const int p = (idx & 7) + 8 * ((idx >> 5) % (TP / 8));
    const int e = ((idx >> 3) & 3) + 4 * (idx / (4 * TP));
    if (ip0 + p >= a.n[a.pd]) continue;
    const long base = (long)(ip0 + p) * st_pd + (long)o * st_od;
    const int t0 = tof(2 * e, N, makhoul), t1 = tof(2 * e + 1, N, makhoul);
    if (in) {
      double x0 = a.inp[0][base + t0 * st_ax], x1 = a.inp[0][base + t1 * st_ax];
      if (sine && (t0 & 1)) x0 = -x0;
      if (sine && (t1 & 1)) x1 = -x1;
      Z[e * TP + p] = make_double2(x0, x1);
    } else {
      const double2 z = Z[e * TP + p];
      a.outp[0][base + t0 * st_ax] = z.x;
      a.outp[0][base + t1 * st_ax] = z.y;
    }
  }
}

}  // namespace radix
}  // namespace snac
