// Real-to-real transforms along the unit-stride direction of both arrays ("row" transforms), included by fft.cu.
//
// Used by the direct plane solver (direct.cu), which transforms the slab in place along its first plane
// direction.  The register-radix kernels (fft_radix.cuh) put consecutive lanes on consecutive pencils, which is
// the coalesced direction when the transform axis is strided; here the axis itself is contiguous, so a warp owns
// one row: m = n/2 = 32 R complex elements z_j = x_2j + i x_2j+1, R of them per lane (j = 32 r + lane: every load
// instruction of the warp reads 512 contiguous bytes, all R loads in flight at once),
//
//   forward :  radix-R DFT over r in registers, twiddle, 32-point DFT across the lanes by butterfly shuffles
//              (decimation in frequency: lane L ends up with index k2 = bitreverse5(L)), the partner Z[m-k] of the
//              untangling step by one more shuffle (lane 31-L, register R-k1), results transposed through a padded
//              per-warp shared-memory row so that the store is coalesced again;
//   backward:  row staged in shared memory, inverse untangling per lane, decimation-in-time shuffles, conjugate
//              twiddle, inverse radix-R DFT, coalesced store straight from registers.
// Same mathematics and tables as r2r_fast_kernel / fft_radix.cuh (FFTW's R2HC/HC2R halfcomplex order;
// REDFT10/01 and RODFT10/01 through Makhoul's permutation, which costs one more pass through the shared-memory
// row on the permuted side).  Lengths n = 64 R, R = 2, 4, 8, 16 (128 ... 1024).
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
constexpr int WARPS = SNAC_ROW_WARPS;   // per CTA; each warp owns two padded row buffers
__host__ __device__ constexpr int padded(int k) { return k + (k >> 4); }   // conflict-free scattered access, see above

__device__ __forceinline__ Cx shfl_xor(Cx v, int mask) {
  return Cx{__shfl_xor_sync(0xffffffffu, v.x, mask), __shfl_xor_sync(0xffffffffu, v.y, mask)};
}
__device__ __forceinline__ Cx shfl_idx(Cx v, int src) {
  return Cx{__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}
__device__ __forceinline__ int brev5(int l) { return (int)(__brev((unsigned)l) >> 27); }

// e^{-2 pi i k / n} for 0 <= k < m from the table of k <= m/2
__device__ __forceinline__ Cx wn_at(const double *wn, int k, int m) {
  if (2 * k <= m) return Cx{__ldg(wn + 2 * k), __ldg(wn + 2 * k + 1)};
  const int kk = m - k;
  return Cx{-__ldg(wn + 2 * kk), __ldg(wn + 2 * kk + 1)};
}

__device__ __forceinline__ void cp_async8(double *dst, const double *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Warps are persistent: each walks over its rows (row = warp + t * warps of the grid) with the NEXT row already
// on its way into the second shared-memory buffer (cp.async, 8 bytes per element so that it lands in the padded
// layout) while the current one is transformed.  The buffer of the current row doubles as the staging area of its
// output.  KIND: 0 halfcomplex (R2HC / HC2R), 1 cosine (REDFT10 / REDFT01), 2 sine (RODFT10 / RODFT01) -- a
// template parameter, so that the permutation and twiddle code of the other kinds is not there at all (as runtime
// flags they doubled the instruction count of the halfcomplex kernel through predication).
// The butterflies of the shuffle network are written without selects: with sg = -1 on the upper lane of a pair and
// +1 on the lower one, and wq = the stage twiddle on the upper lane and 1 on the lower one,
//   decimation in frequency:  v <- (p + sg v) wq          (lower: v + p, upper: (p - v) w)
//   decimation in time     :  t = v conj(wq), v <- p + sg t (lower: v + w b, upper: a - w v)
// which is 6 FP64 instructions per butterfly for every lane instead of both branches plus four selects.
template <int R, bool FWD, int KIND>
__global__ void __launch_bounds__(WARPS * 32, SNAC_ROW_MINB) r2r_row_kernel(const R2RArgs a, long nrows) {
  constexpr int m = 32 * R, n = 2 * m, PN = padded(n);
  constexpr bool makhoul = KIND != 0, sine = KIND == 2;
  extern __shared__ double row_smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long nwarp = (long)gridDim.x * WARPS;
  long row = (long)blockIdx.x * WARPS + wib;
  if (row >= nrows) return;
  double *buf = row_smem + (size_t)wib * 2 * PN;
  const double *tw = a.tab;                  // e^{-2 pi i q/m}, q < m/2
  const double *wn = a.tab + 2 * (m / 2);    // e^{-2 pi i k/n}, k <= m/2
  const double *wd = wn + 2 * (m / 2 + 1);   // e^{-i pi k/(2n)}, k <= m
  const int k2 = brev5(lane);
  // per stage: the twiddle W_{2h}^{lane mod h} = W_m^{(lane mod h) m / 2h} on the upper lane of a pair (1 on the lower), sign
  Cx wq[5];
  double sg[5];
#pragma unroll
  for (int st = 0; st < 5; ++st) {
    const int h = 16 >> st;
    const bool up = (lane & h) != 0;
    const Cx w = radix::tw_m(tw, (lane & (h - 1)) * (m / (2 * h)), m);
    wq[st] = up ? w : Cx{1.0, 0.0};
    sg[st] = up ? -1.0 : 1.0;
  }
  const int src0 = brev5((32 - k2) & 31);
  const double sc = a.scale;
  // padded position of index k = k1 + R k2 (k1 < R <= 16 never carries into the padding term): pk + k1
  const int pk = padded(R * k2);
  auto fetch = [&](long rw, double *dst) {   // row rw -> dst (padded), asynchronously
    const double *in = a.inp[0] + (rw % a.n[1]) * a.ist[1] + (rw / a.n[1]) * a.ist[2];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int e = 2 * (32 * r + lane);
      double *d = dst + 64 * r + 4 * r + 2 * lane + (lane >> 3);   // padded(e): e + (e >> 4) with e = 64 r + 2 lane
      cp_async8(d, in + e);
      cp_async8(d + 1, in + e + 1);
    }
    cp_commit();
  };
  fetch(row, buf);
  int cur = 0;
  for (; row < nrows; row += nwarp, cur ^= 1) {
    double *s = buf + (size_t)cur * PN;
    double2 *out2 = reinterpret_cast<double2 *>(a.outp[0] + (row % a.n[1]) * a.ost[1] + (row / a.n[1]) * a.ost[2]);
    const bool more = row + nwarp < nrows;
    if (more) fetch(row + nwarp, buf + (size_t)(cur ^ 1) * PN);   // that buffer's stores of the previous row are done (below)
    if (more) cp_wait<1>(); else cp_wait<0>();
    __syncwarp();
    const double *sl = s + 2 * lane + (lane >> 3);   // padded(64 r + 2 lane) = sl + 68 r
    Cx v[R];
    if (FWD) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (!makhoul) {
          v[r] = Cx{sl[68 * r], sl[68 * r + 1]};
        } else {
          const int e = 32 * r + lane;
          const int t0 = radix::tof(2 * e, n, true), t1 = radix::tof(2 * e + 1, n, true);
          double x0 = s[padded(t0)], x1 = s[padded(t1)];
          if (sine && (t0 & 1)) x0 = -x0;
          if (sine && (t1 & 1)) x1 = -x1;
          v[r] = Cx{x0, x1};
        }
      }
      __syncwarp();   // every lane has its inputs: the buffer becomes the staging area of the outputs
      radix::dft<R, false>(v);
#pragma unroll
      for (int k1 = 1; k1 < R; ++k1) v[k1] = cmul(v[k1], radix::tw_m(tw, lane * k1, m));
#pragma unroll
      for (int st = 0; st < 5; ++st) {
        const int h = 16 >> st;
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
          const Cx p = shfl_xor(v[k1], h);
          const Cx d{fma(sg[st], v[k1].x, p.x), fma(sg[st], v[k1].y, p.y)};
          v[k1] = cmul(d, wq[st]);
        }
      }
      // v[k1] = Z[k1 + R k2]; its partner Z[m - k] sits in register (R - k1) mod R of lane 31 - lane
      // (k1 > 0) or of the lane holding k2' = (32 - k2) mod 32 (k1 = 0)
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        const Cx zm = k1 == 0 ? shfl_idx(v[0], src0) : shfl_idx(v[(R - k1) % R], 31 - lane);
        const Cx zk = v[k1];
        const int k = k1 + R * k2;
        const Cx e{0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y)};
        const Cx o{0.5 * (zk.y + zm.y), -0.5 * (zk.x - zm.x)};   // -i/2 (zk - conj zm)
        const Cx wo = cmul(wn_at(wn, k, m), o);
        const Cx xk{e.x + wo.x, e.y + wo.y};                       // X[k]
        if (!makhoul) {
          s[pk + k1] = sc * xk.x;
          if (k > 0) s[padded(n - k)] = sc * xk.y;
          else s[padded(m)] = sc * (e.x - wo.x);                   // X[m] (real)
        } else {
          // X_c[k] = 2 Re(e^{-i pi k/2n} V[k]), X_c[n-k] = -2 Im(e^{-i pi k/2n} V[k])
          const double c = __ldg(wd + 2 * k), s_ = -__ldg(wd + 2 * k + 1);
          s[padded(sine ? n - 1 - k : k)] = sc * 2.0 * (c * xk.x + s_ * xk.y);
          if (k > 0) {
            s[padded(sine ? k - 1 : n - k)] = sc * 2.0 * (s_ * xk.x - c * xk.y);
          } else {
            const double cm = __ldg(wd + 2 * m), sm = -__ldg(wd + 2 * m + 1);
            const double xmx = e.x - wo.x, xmy = -(e.y - wo.y);
            s[padded(sine ? n - 1 - m : m)] = sc * 2.0 * (cm * xmx + sm * xmy);
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int r = 0; r < R; ++r) out2[32 * r + lane] = make_double2(sl[68 * r], sl[68 * r + 1]);
    } else {
#pragma unroll
      for (int k1 = 0; k1 < R; ++k1) {
        const int k = k1 + R * k2, kb = m - k;
        Cx xk, xm;   // X[k], X[m-k]
        if (!makhoul) {
          xk.x = s[pk + k1];    xk.y = k > 0 ? s[padded(n - k)] : 0.0;
          xm.x = s[padded(kb)]; xm.y = kb < m ? s[padded(n - kb)] : 0.0;
        } else {
          auto getx = [&](int kk) -> Cx {
            const double c = __ldg(wd + 2 * kk), s_ = -__ldg(wd + 2 * kk + 1);
            const double a0 = s[padded(sine ? n - 1 - kk : kk)];
            const double a1 = kk > 0 ? s[padded(sine ? kk - 1 : n - kk)] : 0.0;
            return Cx{c * a0 + s_ * a1, s_ * a0 - c * a1};
          };
          xk = getx(k);
          xm = getx(kb);
        }
        // Z[k] = (X[k] + conj X[m-k]) + i e^{+2 pi i k/n} (X[k] - conj X[m-k])
        Cx w = wn_at(wn, k, m);
        w.y = -w.y;
        const Cx sum{xk.x + xm.x, xk.y - xm.y}, dif{xk.x - xm.x, xk.y + xm.y};
        const Cx wd_ = cmul(w, dif);
        v[k1] = Cx{sum.x - wd_.y, sum.y + wd_.x};
      }
      // inverse 32-point DFT across the lanes: decimation in time, bit-reversed in, natural out
#pragma unroll
      for (int st = 4; st >= 0; --st) {
        const int h = 16 >> st;
        const Cx wc{wq[st].x, -wq[st].y};
#pragma unroll
        for (int k1 = 0; k1 < R; ++k1) {
          const Cx t = cmul(v[k1], wc);
          const Cx p = shfl_xor(t, h);
          v[k1] = Cx{fma(sg[st], t.x, p.x), fma(sg[st], t.y, p.y)};
        }
      }
#pragma unroll
      for (int k1 = 1; k1 < R; ++k1) {
        Cx w = radix::tw_m(tw, lane * k1, m);
        w.y = -w.y;
        v[k1] = cmul(v[k1], w);
      }
      radix::dft<R, true>(v);
      if (!makhoul) {
#pragma unroll
        for (int r = 0; r < R; ++r) out2[32 * r + lane] = make_double2(v[r].x * sc, v[r].y * sc);
      } else {
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int e = 32 * r + lane;
          const int t0 = radix::tof(2 * e, n, true), t1 = radix::tof(2 * e + 1, n, true);
          double x0 = v[r].x * sc, x1 = v[r].y * sc;
          if (sine && (t0 & 1)) x0 = -x0;
          if (sine && (t1 & 1)) x1 = -x1;
          s[padded(t0)] = x0; s[padded(t1)] = x1;
        }
        __syncwarp();
#pragma unroll
        for (int r = 0; r < R; ++r) out2[32 * r + lane] = make_double2(sl[68 * r], sl[68 * r + 1]);
      }
    }
    __syncwarp();   // the buffer is free again: it receives the row after the next one
  }
}

template <int R, bool FWD, int KIND>
static void row_go(cudaStream_t st, const R2RArgs &a, long nrows) {
  constexpr int n = 64 * R;
  const size_t smem = (size_t)WARPS * 2 * padded(n) * sizeof(double);
  static bool attr = false;
  static int per_sm = 1, sms = 148;
  if (!attr) {
    if (smem > 48 * 1024)
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
