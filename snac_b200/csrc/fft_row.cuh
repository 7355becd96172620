// Real-to-real transforms along the unit-stride direction of both arrays ("row" transforms), included by fft.cu.
//
// Used by the direct plane solver (direct.cu), which transforms the slab in place along its first plane
// direction.  The register-radix kernels (fft_radix.cuh) put consecutive lanes on consecutive pencils, which is
// the coalesced direction when the transform axis is strided; here the axis itself is contiguous, so a warp owns
// one row.  m = n/2 = 32 R complex elements z_j = x_2j + i x_2j+1 are transformed as m = R * R * G, G = 32 / R:
//
//   A  lane l holds z[32 r + l], r < R: radix-R DFT over r in registers, twiddle W_m^(l k1)
//      -> transposed through shared memory: lane (k1, g) = k1 G + g takes the R values l = g + G t of its k1
//   B  radix-R DFT over t in registers, twiddle W_32^(g u)
//   C  DFT of length G across the G adjacent lanes of a k1 by log2 G butterfly shuffles (none needed but one
//      exchange for n = 1024): lane (k1, g), register u ends with Z[k1 + R u + R^2 bitrev_G(g)]
//   -> written to shared memory in natural order; the untangling step X[k] = E[k] + W_n^k O[k] then runs over
//      consecutive k per lane (Z[k] and Z[m-k] are two conflict-free shared-memory reads) and stores straight to
//      global memory, r_k at k and i_k at n-k: both coalesced.
// The backward transform is the same machine: the inverse untangling produces conj(Z[k]) for the lane's own
// k = l + 32 r -- exactly the registers stage A starts from -- and the result is conj(.) of the forward output.
// The first version of this kernel did the 32-point cross-lane DFT with five shuffle stages (384 shuffles per lane
// and row) and fetched the untangling partner by shuffles too; ncu showed it bound by the shuffle/shared-memory
// instruction queue (mio_throttle, short_scoreboard; FP64 pipe 24 % busy, 2 TB/s).  This form issues a third of
// those instructions; twiddles that depend on the lane are rebuilt from one per-lane base by multiplication
// (p_k = p_(k/2) p_((k+1)/2)) instead of table loads, the others are compile-time constants.
// Same mathematics and output conventions as r2r_fast_kernel / fft_radix.cuh (FFTW's R2HC/HC2R halfcomplex
// order; REDFT10/01 and RODFT10/01 through Makhoul's permutation).  Lengths n = 64 R, R = 2, 4, 8, 16.
// Warps are persistent; the next row is fetched by cp.async into the warp's input buffer as soon as stage A has
// taken the current one out of it.
#pragma once
#include "fft_radix.cuh"

namespace snac {
namespace rowfft {

using radix::Cx;
using radix::cmul;

#ifndef SNAC_ROW_WARPS
#define SNAC_ROW_WARPS 4
#endif
#ifndef SNAC_ROW_MINB
#define SNAC_ROW_MINB 3
#endif
constexpr int WARPS = SNAC_ROW_WARPS;   // per CTA; each warp owns an input row and a scratch area

// cos, sin of 2 pi k / 32 and 2 pi k / 128, k < 16
#define SNAC_C32 {1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476, 0.5555702330196023, \
                  0.38268343236508984, 0.19509032201612833, 0.0, -0.1950903220161282, -0.3826834323650897, -0.555570233019602, \
                  -0.7071067811865475, -0.8314696123025453, -0.9238795325112867, -0.9807852804032304}
#define SNAC_S32 {0.0, 0.19509032201612825, 0.3826834323650898, 0.5555702330196022, 0.7071067811865475, 0.8314696123025452, \
                  0.9238795325112867, 0.9807852804032304, 1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025455, \
                  0.7071067811865476, 0.5555702330196022, 0.3826834323650899, 0.1950903220161286}
#define SNAC_C128 {1.0, 0.9987954562051724, 0.9951847266721969, 0.989176509964781, 0.9807852804032304, 0.970031253194544, \
                   0.9569403357322088, 0.9415440651830208, 0.9238795325112867, 0.9039892931234433, 0.881921264348355, \
                   0.8577286100002721, 0.8314696123025452, 0.8032075314806449, 0.773010453362737, 0.7409511253549591}
#define SNAC_S128 {0.0, 0.049067674327418015, 0.0980171403295606, 0.14673047445536175, 0.19509032201612825, \
                   0.24298017990326387, 0.29028467725446233, 0.33688985339222005, 0.3826834323650898, 0.4275550934302821, \
                   0.47139673682599764, 0.5141027441932217, 0.5555702330196022, 0.5956993044924334, 0.6343932841636455, \
                   0.6715589548470183}

__device__ __forceinline__ Cx shfl_xor(Cx v, int mask) {
  return Cx{__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask)};
}
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// p[k] = t^k, k < R, by products of two lower powers (error grows with log k)
template <int R>
__device__ __forceinline__ void powers(Cx t, Cx (&p)[R]) {
  p[0] = Cx{1.0, 0.0};
  if (R > 1) p[1] = t;
#pragma unroll
  for (int k = 2; k < R; ++k) p[k] = cmul(p[k >> 1], p[(k + 1) >> 1]);
}

template <int R>
struct LaneConst {       // what a lane keeps across its rows
  Cx tA, tB;             // W_m^lane, W_32^g
  Cx wq[5];              // stage-C twiddles (1 on the lower lane of a pair)
  double sg[5];          // -1 on the upper lane, +1 on the lower one
  int k1p, g, sB;
};

// stages A, B, C: v[r] = z[32 r + lane] on entry; on exit v[u] = Z[k1p + R u + R^2 sB].  scr: R (32 + G) complex.
template <int R>
__device__ __forceinline__ void core(Cx (&v)[R], double2 *scr, const LaneConst<R> &c, int lane) {
  constexpr int G = 32 / R, LG = radix::log2_c(G), TS = 32 + G;
  radix::dft<R, false>(v);
  {
    Cx p[R];
    powers<R>(c.tA, p);
#pragma unroll
    for (int k1 = 1; k1 < R; ++k1) v[k1] = cmul(v[k1], p[k1]);
  }
#pragma unroll
  for (int k1 = 0; k1 < R; ++k1) scr[k1 * TS + lane] = make_double2(v[k1].x, v[k1].y);
  __syncwarp();
#pragma unroll
  for (int t = 0; t < R; ++t) {
    const double2 z = scr[c.k1p * TS + c.g + G * t];
    v[t] = Cx{z.x, z.y};
  }
  radix::dft<R, false>(v);
  if (G > 1) {
    Cx p[R];
    powers<R>(c.tB, p);
#pragma unroll
    for (int u = 1; u < R; ++u) v[u] = cmul(v[u], p[u]);
  }
#pragma unroll
  for (int st = 0; st < LG; ++st) {
    const int h = G >> (st + 1);
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const Cx p = shfl_xor(v[u], h);
      const Cx d{fma(c.sg[st], v[u].x, p.x), fma(c.sg[st], v[u].y, p.y)};
      v[u] = h > 1 ? cmul(d, c.wq[st]) : d;   // the last stage has no twiddle
    }
  }
  __syncwarp();   // every lane has read its column of the transpose: the scratch area is free
}

// KIND: 0 halfcomplex (R2HC / HC2R), 1 cosine (REDFT10 / REDFT01), 2 sine (RODFT10 / RODFT01)
template <int R, bool FWD, int KIND>
__global__ void __launch_bounds__(WARPS * 32, SNAC_ROW_MINB) r2r_row_kernel(const R2RArgs a, long nrows) {
  constexpr int G = 32 / R, LG = radix::log2_c(G), m = 32 * R, n = 2 * m, TS = 32 + G;
  constexpr int WSM = n + 2 * R * TS;     // doubles per warp: input row + scratch
  constexpr bool makhoul = KIND != 0, sine = KIND == 2;
  constexpr double c32[16] = SNAC_C32, s32[16] = SNAC_S32, c128[16] = SNAC_C128, s128[16] = SNAC_S128;
  extern __shared__ __align__(16) double row_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long nwarp = (long)gridDim.x * WARPS;
  long row = (long)blockIdx.x * WARPS + wib;
  if (row >= nrows) return;
  double *in = row_smem + (size_t)wib * WSM;
  double2 *in2 = reinterpret_cast<double2 *>(in);
  double2 *scr = reinterpret_cast<double2 *>(in + n);
  double *scd = in + n;
  const double *tw = a.tab;                  // e^{-2 pi i q/m}, q < m/2
  const double *wn = a.tab + 2 * (m / 2);    // e^{-2 pi i k/n}, k <= m/2
  const double *wd = wn + 2 * (m / 2 + 1);   // e^{-i pi k/(2n)}, k <= m
  LaneConst<R> c;
  c.k1p = lane / G; c.g = lane % G;
  c.sB = (int)(__brev((unsigned)c.g) >> (32 - LG)) & (G - 1);
  if (LG == 0) c.sB = 0;
  c.tA = radix::tw_m(tw, lane, m);
  c.tB = radix::tw_m(tw, c.g * R, m);
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    c.wq[st] = Cx{1.0, 0.0}; c.sg[st] = 1.0;
    if (st < LG) {
      const int h = G >> (st + 1);
      const bool up = (c.g & h) != 0;
      if (up) { c.wq[st] = radix::tw_m(tw, (c.g & (h - 1)) * (m / (2 * h)), m); c.sg[st] = -1.0; }
    }
  }
  const Cx wnl{__ldg(wn + 2 * lane), __ldg(wn + 2 * lane + 1)};   // W_n^lane   (lane <= 31 <= m/2)
  const Cx wdl{__ldg(wd + 2 * lane), __ldg(wd + 2 * lane + 1)};   // e^{-i pi lane/(2n)}
  const double sc = a.scale;
  auto fetch = [&](long rw) {   // row rw -> the warp's input buffer, asynchronously
    const double *src = a.inp[0] + (rw % a.n[1]) * a.ist[1] + (rw / a.n[1]) * a.ist[2];
#pragma unroll
    for (int r = 0; r < R; ++r) cp_async16(in2 + 32 * r + lane, src + 2 * (32 * r + lane));
    cp_commit();
  };
  fetch(row);
#pragma unroll 1
  for (; row < nrows; row += nwarp) {
    double *out = a.outp[0] + (row % a.n[1]) * a.ost[1] + (row / a.n[1]) * a.ost[2];
    cp_wait_all();
    __syncwarp();
    Cx v[R];
    if (FWD) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!makhoul) {
          const double2 z = in2[32 * r + lane];
          v[r] = Cx{z.x, z.y};
        } else {
          const int e = 32 * r + lane;
          const int t0 = radix::tof(2 * e, n, true), t1 = radix::tof(2 * e + 1, n, true);
          double x0 = in[t0], x1 = in[t1];
          if (sine && (t0 & 1)) x0 = -x0;
          if (sine && (t1 & 1)) x1 = -x1;
          v[r] = Cx{x0, x1};
        }
      }
    } else {
      // inverse untangling for the lane's own k = lane + 32 r: conj(Z[k]),
      // Z[k] = (X[k] + conj X[m-k]) + i e^{+2 pi i k/n} (X[k] - conj X[m-k])
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = lane + 32 * r, kb = m - k;
        const int j = r * (16 / R);   // 32 r / n and 32 r / 2n as fractions of the constant tables
        Cx xk, xm;
        if (!makhoul) {
          xk.x = in[k];  xk.y = k > 0 ? in[n - k] : 0.0;
          xm.x = in[kb]; xm.y = kb < m ? in[m + k] : 0.0;
        } else {
          // e^{-i pi k/2n} = wdl e^{-2 pi i j/128};  e^{-i pi (m-k)/2n} = e^{-i pi/4} conj(e^{-i pi k/2n})
          const Cx wk = cmul(wdl, Cx{c128[j], -s128[j]});
          const Cx wb = cmul(Cx{0.70710678118654752440, -0.70710678118654752440}, Cx{wk.x, -wk.y});
          auto getx = [&](int kk, Cx w) -> Cx {
            const double cc = w.x, s_ = -w.y;
            const double a0 = in[sine ? n - 1 - kk : kk];
            const double a1 = kk > 0 ? in[sine ? kk - 1 : n - kk] : 0.0;
            return Cx{cc * a0 + s_ * a1, s_ * a0 - cc * a1};
          };
          xk = getx(k, wk);
          xm = getx(kb, wb);
        }
        const Cx wf = cmul(wnl, Cx{c32[j], -s32[j]});   // W_n^k
        const Cx w{wf.x, -wf.y};
        const Cx sum{xk.x + xm.x, xk.y - xm.y}, dif{xk.x - xm.x, xk.y + xm.y};
        const Cx wd_ = cmul(w, dif);
        v[r] = Cx{sum.x - wd_.y, -(sum.y + wd_.x)};
      }
    }
    __syncwarp();                                   // every lane has taken its part of the row out of the buffer
    if (row + nwarp < nrows) fetch(row + nwarp);    // ... which receives the next one while this one is transformed
    core<R>(v, scr, c, lane);
    if (FWD) {
#pragma unroll
      for (int u = 0; u < R; ++u) scr[c.k1p + R * u + R * R * c.sB] = make_double2(v[u].x, v[u].y);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int k = lane + 32 * r;
        const int j = r * (16 / R);
        const double2 zk2 = scr[k], zm2 = scr[(m - k) & (m - 1)];
        const Cx e{0.5 * (zk2.x + zm2.x), 0.5 * (zk2.y - zm2.y)};
        const Cx o{0.5 * (zk2.y + zm2.y), -0.5 * (zk2.x - zm2.x)};   // -i/2 (zk - conj zm)
        const Cx wo = cmul(cmul(wnl, Cx{c32[j], -s32[j]}), o);
        const Cx xk{e.x + wo.x, e.y + wo.y};                           // X[k]
        if (!makhoul) {
          out[k] = sc * xk.x;
          if (k > 0) out[n - k] = sc * xk.y;
          else out[m] = sc * (e.x - wo.x);                             // X[m] (real)
        } else {
          // X_c[k] = 2 Re(e^{-i pi k/2n} V[k]), X_c[n-k] = -2 Im(e^{-i pi k/2n} V[k])
          const Cx wk = cmul(wdl, Cx{c128[j], -s128[j]});
          const double cc = wk.x, s_ = -wk.y;
          out[sine ? n - 1 - k : k] = sc * 2.0 * (cc * xk.x + s_ * xk.y);
          if (k > 0) {
            out[sine ? k - 1 : n - k] = sc * 2.0 * (s_ * xk.x - cc * xk.y);
          } else {
            const double cm = 0.70710678118654752440, sm = 0.70710678118654752440;   // e^{-i pi/4}: cos, -(-sin)
            const double xmx = e.x - wo.x, xmy = -(e.y - wo.y);
            out[sine ? n - 1 - m : m] = sc * 2.0 * (cm * xmx + sm * xmy);
          }
        }
      }
    } else {
      // v[u] = conj(z[j]), j = k1p + R u + R^2 sB: to natural order through the scratch area, then a coalesced store
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int j = c.k1p + R * u + R * R * c.sB;
        double x0 = v[u].x * sc, x1 = -v[u].y * sc;
        if (!makhoul) {
          scr[j] = make_double2(x0, x1);
        } else {
          const int t0 = radix::tof(2 * j, n, true), t1 = radix::tof(2 * j + 1, n, true);
          if (sine && (t0 & 1)) x0 = -x0;
          if (sine && (t1 & 1)) x1 = -x1;
          scd[t0] = x0; scd[t1] = x1;
        }
      }
      __syncwarp();
      double2 *out2 = reinterpret_cast<double2 *>(out);
#pragma unroll
      for (int r = 0; r < R; ++r) out2[32 * r + lane] = scr[32 * r + lane];
    }
    __syncwarp();   // the scratch area is free for the next row
  }
}

template <int R, bool FWD, int KIND>
static void row_go(cudaStream_t st, const R2RArgs &a, long nrows) {
  constexpr int n = 64 * R, G = 32 / R;
  const size_t smem = (size_t)WARPS * (n + 2 * R * (32 + G)) * sizeof(double);
  static bool attr = false;
  static int per_sm = 1, sms = 148;
  if (!attr) {
    if (smem > 40 * 1024)
      SNAC_CUDA(cudaFuncSetAttribute(r2r_row_kernel<R, FWD, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, r2r_row_kernel<R, FWD, KIND>, WARPS * 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    attr = true;
  }
  const long ctas = std::min<long>((nrows + WARPS - 1) / WARPS, (long)sms * per_sm);
  r2r_row_kernel<R, FWD, KIND><<<(unsigned)ctas, WARPS * 32, smem, st>>>(a, nrows);
}

template <int R>
static void row_kind(cudaStream_t st, const R2RArgs &a, int dir, long nrows) {
  const int kind = radix::is_makhoul(a.kind) ? (radix::is_sine(a.kind) ? 2 : 1) : 0;
  if (dir > 0) {
    if (kind == 0) row_go<R, true, 0>(st, a, nrows); else if (kind == 1) row_go<R, true, 1>(st, a, nrows); else row_go<R, true, 2>(st, a, nrows);
  } else {
    if (kind == 0) row_go<R, false, 0>(st, a, nrows); else if (kind == 1) row_go<R, false, 1>(st, a, nrows); else row_go<R, false, 2>(st, a, nrows);
  }
}

// the transform axis is x (unit stride in both arrays), one array on either side, 16-byte aligned rows
static bool row_launch(cudaStream_t st, const R2RArgs &a, int dir) {
  if (a.axis != 0 || a.ist[0] != 1 || a.ost[0] != 1 || a.in_split || a.out_split) return false;
  if (a.kind == K_REDFT11 || a.kind == K_RODFT11 || a.kind == K_REDFT00 || a.kind == K_RODFT00 || a.kind == K_DHT) return false;
  if ((reinterpret_cast<size_t>(a.inp[0]) | reinterpret_cast<size_t>(a.outp[0])) & 15) return false;
  if ((a.ist[1] | a.ist[2] | a.ost[1] | a.ost[2]) & 1) return false;
  const long nrows = (long)a.n[1] * a.n[2];
  switch (a.nt) {
    case 128:  row_kind<2>(st, a, dir, nrows);  return true;
    case 256:  row_kind<4>(st, a, dir, nrows);  return true;
    case 512:  row_kind<8>(st, a, dir, nrows);  return true;
    case 1024: row_kind<16>(st, a, dir, nrows); return true;
    default: return false;
  }
}

}  // namespace rowfft
}  // namespace snac
