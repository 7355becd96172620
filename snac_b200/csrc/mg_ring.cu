// K4, plane-local fast path: the two fused fine-level kernels of a V(1,1) cycle in their
// "prefetch ring" form (see mg.cu for the operator, the cycle and the reference call sites,
// src/solver.f90:843-922).
//
//   down: red sweep + black sweep + residual + restriction      (read x, b; write x', coarse b)
//   up  : coarse-grid correction + black sweep + red sweep      (read x', b, coarse x; write x)
//
// A warp owns a strip of 64 columns, two adjacent ones (an aligned 16-byte pair = the two children
// of one coarse cell) per lane, and marches over the rows of its chunk.  The rows of x and b, the
// coarse correction and the row's operator coefficients are streamed from HBM into a per-warp
// shared-memory ring by cp.async RP rows ahead of their use: the ring is the deep, register-free
// prefetch queue that keeps enough bytes in flight per SM to cover the HBM latency.  Each lane
// copies and later reads back only its own 16 bytes of x and b (no barrier, no cross-lane
// visibility needed); the 64-byte coefficient entry of a row is copied by lanes 0-3 and read by
// all after a __syncwarp.  The stencil works on a register window of six rows of x; horizontal
// neighbours come from the lane's other column or one warp shuffle away; b and the coefficients
// are read from the ring slot of their row when they are used (a slot lives RD - RP - 1 steps
// after its row entered the window).  At step t: row t+1 enters the window, the first colour of
// row t, the second colour of row t-1 and the output of row t-2 (down: store, residual, its norm
// and its restriction; up: store of row t-1) are done.  Strips overlap by HX columns and chunks by
// HY rows on each side, which are recomputed; a periodic level that is exactly 64 columns wide
// needs no column halo (the shuffles wrap around).
//
// Cell equations are used in the scaled form (division by w1*w2)
//   x = ( b/(w1 w2) + na1m xL + na1p xR + na2m xD + na2p xU ) / (e1 + e2 + lambda),  na = -am/w, e = ad/w,
// with packed per-direction tables {na_m, na_p, e, w, 1/w, 0, 0, 0} built on the host (entry 0 = a
// cell outside a non-periodic domain: na = 0, e = -1, w = 1/w = 0, so that its update yields
// exactly 0 without a branch and nothing propagates from it).  1/(e1+e2+lambda) uses the hardware
// reciprocal seed + two Newton steps (full double precision, no slow path: the diagonal is never
// zero, subnormal or infinite here).
#include "common.h"

#include <algorithm>
#include <cstdlib>

namespace snac {

namespace {
constexpr int RW = 64;    // columns per warp
#ifndef SNAC_RING_RD
#define SNAC_RING_RD 8
#endif
#ifndef SNAC_RING_RPD
#define SNAC_RING_RPD 2
#endif
// ring slots = window slots = unroll period of the march, down kernel.  Measured on TGV-256 (B200, same
// box): 12 slots / 6 rows ahead (4 CTAs per SM) 2.37 ms per step, 10 / 4 (5 CTAs) 1.76-1.86 ms, 8 / 2
// (5 CTAs, shorter unrolled body) 1.68 ms; 8 / 2 squeezed into 80 registers for 6 CTAs: 2.02 ms.
constexpr int RD = SNAC_RING_RD;
#ifndef SNAC_RING_RDU
#define SNAC_RING_RDU 8
#endif
#ifndef SNAC_RING_MINB_D
#define SNAC_RING_MINB_D 5
#endif
#ifndef SNAC_RING_MINB_U
#define SNAC_RING_MINB_U 1
#endif
constexpr int RDU = SNAC_RING_RDU;    // the same for the up kernel (rows leave the window earlier there)
// rows in flight ahead of the one entering the window; a slot is refilled at the end of step t with row
// t + 1 + RP_D, i.e. it must hold a row last used at step t - 1 or earlier: RD - RP_D - 1 >= DOUT + 1
constexpr int RP_D = SNAC_RING_RPD;
static_assert(RD % 2 == 0 && RD - RP_D - 1 >= 5 && RP_D >= 2 && RP_D <= 6 && (RD == 8 || RD == 10 || RD == 12), "ring geometry");
static_assert(RDU == 6 || RDU == 8, "ring geometry");
constexpr int WARPS_D = 4, WARPS_U = 2;   // warps per CTA (independent of each other)
constexpr int SLOT_D = 64 + 64 + 8;       // doubles per ring slot, down: x, b, coefficients
constexpr int SLOT_U = 64 + 64 + 64 + 8;  // up: x, b, coarse x, coefficients

__device__ __forceinline__ double fast_rcp(double e) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(e));
  double t = fma(-e, r, 1.0);
  r = fma(r, t, r);
  t = fma(-e, r, 1.0);
  return fma(r, t, r);
}
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct ColPair {   // the lane's two columns
  double nam0, nap0, e0, w0, rw0, nam1, nap1, e1, w1, rw1;
  int gi0;         // memory column of the pair (odd; 1 when there is nothing to load, so that addresses stay valid)
  int ld;          // bytes of the pair that exist (0, 8, 16)
  int st;          // bytes of the pair this warp keeps (0, 8, 16)
};

// WRAP: the strip is the whole periodic row (n1 == 64), no column halo
template <int HX, bool WRAP>
__device__ __forceinline__ void ring_cols(const RLevel &L, int i0, double lam, int lane, ColPair &C) {
  const int gu0 = WRAP ? 2 * lane + 1 : i0 - HX + 2 * lane;   // odd
  int g0 = gu0, g1 = gu0 + 1;
  if (!WRAP) {
    if (L.per1) {
      if (g0 < 1) { g0 += L.n1; g1 += L.n1; }
      else if (g0 > L.n1) { g0 -= L.n1; g1 -= L.n1; }
      if (gu0 > L.n1 + HX) { g0 = 0; g1 = 0; }     // overhang of the last strip: never needed
    } else {
      if (g0 < 1 || g0 > L.n1) g0 = 0;
      if (g1 < 1 || g1 > L.n1) g1 = 0;
    }
  }
  const double *t0 = L.ctab + 8 * (long)g0, *t1 = L.ctab + 8 * (long)g1;
  C.nam0 = __ldg(t0); C.nap0 = __ldg(t0 + 1); C.e0 = __ldg(t0 + 2) + (g0 ? lam : 0.0); C.w0 = __ldg(t0 + 3); C.rw0 = __ldg(t0 + 4);
  C.nam1 = __ldg(t1); C.nap1 = __ldg(t1 + 1); C.e1 = __ldg(t1 + 2) + (g1 ? lam : 0.0); C.w1 = __ldg(t1 + 3); C.rw1 = __ldg(t1 + 4);
  C.gi0 = g0 ? g0 : 1;
  C.ld = g0 ? (g1 ? 16 : 8) : 0;
  const bool own0 = WRAP || (2 * lane >= HX && 2 * lane < RW - HX && gu0 <= L.n1);
  const bool own1 = WRAP || (2 * lane + 1 >= HX && 2 * lane + 1 < RW - HX && gu0 + 1 <= L.n1);
  C.st = own0 ? (own1 ? 16 : 8) : 0;
}

// memory row of global row gv (1-based, possibly outside): 0 = outside / not needed
template <int HY>
__device__ __forceinline__ int ring_row(const RLevel &L, int gv) {
  const int w = gv < 1 ? gv + L.n2 : (gv > L.n2 ? gv - L.n2 : gv);
  const bool in = L.per2 ? gv <= L.n2 + HY : (gv >= 1 && gv <= L.n2);
  return in ? w : 0;
}

// Gauss-Seidel update of one cell.  A: the cell is in the lane's first column; xs/xo: this row's
// own-column and other-column values; xm/xp: the own column one row below/above; slot: ring slot
// of the row (b pair at [64 + 2 lane], coefficients at [CO])
template <bool WRAP, int CO>
__device__ __forceinline__ double ring_upd(bool A, const ColPair &C, const double *slot, int lane, double xo, double xm, double xp) {
  const double2 na2 = *reinterpret_cast<const double2 *>(slot + CO);        // na2m, na2p
  const double2 ew = *reinterpret_cast<const double2 *>(slot + CO + 2);     // e2, w2
  const double rw2 = slot[CO + 4];
  const double bs = slot[64 + 2 * lane + (A ? 0 : 1)] * ((A ? C.rw0 : C.rw1) * rw2);
  const double l = A ? (WRAP ? __shfl_sync(0xffffffffu, xo, (lane + 31) & 31) : __shfl_up_sync(0xffffffffu, xo, 1)) : xo;
  const double r = A ? xo : (WRAP ? __shfl_sync(0xffffffffu, xo, (lane + 1) & 31) : __shfl_down_sync(0xffffffffu, xo, 1));
  const double inv = fast_rcp((A ? C.e0 : C.e1) + ew.x);
  double acc = fma(A ? C.nam0 : C.nam1, l, bs);
  double acc2 = na2.x * xm;
  acc = fma(A ? C.nap0 : C.nap1, r, acc);
  acc2 = fma(na2.y, xp, acc2);
  return (acc + acc2) * inv;
}

constexpr int CM_FULL = 0, CM_X = 1, CM_Y = 2;   // directions coarsened towards the next level

// ---- down ------------------------------------------------------------------------------------
// NSW Gauss-Seidel sweeps (colour 0 then 1, each), residual, restriction.
//   NSW = 1: colour 0 of row t, colour 1 of row t-2, output of row t-4: the phases of a step work on
//            rows two apart, so that none of them waits for another one of the same step;
//   NSW = 2: colours 0,1,0,1 of rows t, t-1, t-2, t-3, output of row t-4 (a dependent chain inside
//            the step, hidden by the other warps; keeps the lifetime of a ring slot at 4 steps).
template <bool WRAP, bool HASX, int CM, int NSW>
struct Down {
  static constexpr int HY = NSW == 1 ? 3 : 5;                 // rows recomputed above and below a chunk
  static constexpr int HX = WRAP ? 0 : (NSW == 1 ? 4 : 6);    // columns recomputed left and right (even)
  static constexpr int TX = RW - 2 * HX, CO = 128;
  static constexpr int DOUT = 4;                              // the output of a row lags its entry by DOUT steps
  const RLevel &L;
  ColPair C;
  double *ring;            // this warp's ring (generic address)
  unsigned rbase;          // shared-space address of the lane's x pair in slot 0
  unsigned cbase;          // shared-space address of the lane's coefficient chunk in slot 0 (lanes 0-3)
  const double *xsrc, *bsrc;
  double *xdst, *cbq;
  int csy, lane, j0, mry, HT, want_b;
  double wc0, wc1, xscale;   // xscale: factor applied to the incoming iterate (initial guess)
  double xa[RD], xb[RD];   // register window (indexed by row mod RD; only about seven rows are live at a time)
  double r2 = 0.0, b2 = 0.0, resP = 0.0;

  __device__ __forceinline__ Down(const RLevel &L_) : L(L_) {}

  template <int SL>
  __device__ __forceinline__ void issue(int t) {   // copy of tile row t into ring slot SL
    const int gj = ring_row<HY>(L, j0 - HY + t);
    const bool on = gj && t < HT;
    const int nb = on ? C.ld : 0;
    const long o = (long)L.sy * gj;
    if (HASX) cp_async16(rbase + SL * SLOT_D * 8, xsrc + o, nb);
    cp_async16(rbase + (SL * SLOT_D + 64) * 8, bsrc + o, nb);
    if (lane < 4) cp_async16(cbase + SL * SLOT_D * 8, L.rtab + 8 * (on ? gj : 0) + 2 * lane, 16);
    cp_commit();
  }

  // colour COLOR of row t - D (K = t mod RD; t0 is a multiple of RD, which is even).  Row r
  // (tile-local) has its first column in colour 0 iff (r + HY) is even.
  template <int K, int D, int COLOR>
  __device__ __forceinline__ void upd() {
    constexpr int R = (K + RD - D) % RD, Rm = (K + 2 * RD - D - 1) % RD, Rp = (K + RD - D + 1) % RD;
    constexpr bool first0 = ((K + RD - D + HY) & 1) == 0;
    constexpr bool A = (COLOR == 0) == first0;   // the cell of this colour sits in the lane's first column
    const double *slot = ring + R * SLOT_D;
    if (A) xa[R] = ring_upd<WRAP, CO>(true, C, slot, lane, xb[R], xa[Rm], xa[Rp]);
    else xb[R] = ring_upd<WRAP, CO>(false, C, slot, lane, xa[R], xb[Rm], xb[Rp]);
  }

  // one step; K = t mod 12: ring slots and window slots are compile-time constants.  The first
  // steps work on rows that are not there yet (zeros): harmless, they only feed halo rows.
  template <int K>
  __device__ __forceinline__ void step(int t0) {
    const int t = t0 + K;
    {  // row t+1 enters the window
      constexpr int S = (K + 1) % RD;
      cp_wait<RP_D - 1>();
      __syncwarp();
      const double *slot = ring + S * SLOT_D;
      if (HASX) {
        const double2 v = *reinterpret_cast<const double2 *>(slot + 2 * lane);
        xa[S] = v.x * xscale; xb[S] = v.y * xscale;
      } else {
        xa[S] = 0.0; xb[S] = 0.0;
      }
      if (want_b) {
        const int tl = t + 1;
        const bool rin = tl >= HY && tl < HY + mry && (j0 - HY + tl) <= L.n2;
        const double2 w = *reinterpret_cast<const double2 *>(slot + 64 + 2 * lane);
        b2 += rin ? (C.st ? w.x * w.x : 0.0) + (C.st == 16 ? w.y * w.y : 0.0) : 0.0;
      }
    }
    if (NSW == 1) {
      upd<K, 0, 0>();
      upd<K, 2, 1>();
    } else {
      upd<K, 0, 0>();
      upd<K, 1, 1>();
      upd<K, 2, 0>();
      upd<K, 3, 1>();
    }
    {  // output of row t - DOUT
      constexpr int R = (K + RD - DOUT) % RD, Rm = (K + RD - DOUT - 1) % RD, Rp = (K + RD - DOUT + 1) % RD;
      const double *slot = ring + R * SLOT_D;
      const int gv = j0 + t - DOUT - HY;
      const bool rin = t >= HY + DOUT && t < HY + DOUT + mry && gv <= L.n2;
      const long o = (long)L.sy * gv;
      if (rin) {
        if (C.st == 16) *reinterpret_cast<double2 *>(xdst + o) = make_double2(xa[R], xb[R]);
        else if (C.st == 8) xdst[o] = xa[R];
      }
      const double2 na2 = *reinterpret_cast<const double2 *>(slot + CO);
      const double2 ew = *reinterpret_cast<const double2 *>(slot + CO + 2);
      const double rw2 = slot[CO + 4];
      const double w2 = rin ? ew.y : 0.0;
      double res;
      constexpr bool first = ((K + RD - DOUT + HY) & 1) == 0;   // colour 0 sits in the first column
      if (first) {
        const double lft = WRAP ? __shfl_sync(0xffffffffu, xb[R], (lane + 31) & 31) : __shfl_up_sync(0xffffffffu, xb[R], 1);
        const double bs = slot[64 + 2 * lane] * (C.rw0 * rw2);
        res = bs + C.nam0 * lft + C.nap0 * xb[R] + na2.x * xa[Rm] + na2.y * xa[Rp] - (C.e0 + ew.x) * xa[R];
        res *= wc0 * w2;
      } else {
        const double rgt = WRAP ? __shfl_sync(0xffffffffu, xa[R], (lane + 1) & 31) : __shfl_down_sync(0xffffffffu, xa[R], 1);
        const double bs = slot[64 + 2 * lane + 1] * (C.rw1 * rw2);
        res = bs + C.nam1 * xa[R] + C.nap1 * rgt + na2.x * xb[Rm] + na2.y * xb[Rp] - (C.e1 + ew.x) * xb[R];
        res *= wc1 * w2;
      }
      r2 = fma(res, res, r2);
      if (cbq) {
        const bool child2 = (gv & 1) == 0;   // second child row of its coarse row
        if (CM == CM_FULL) {
          if (rin && C.st && (child2 || gv == L.n2)) cbq[(long)csy * ((gv + 1) >> 1)] = child2 ? resP + res : res;
        } else if (CM == CM_X) {
          if (rin && C.st) cbq[(long)csy * gv] = res;
        } else {   // y only: each column's colour-0 residual comes from one of the two child rows
          if (rin && C.st && (child2 || gv == L.n2)) {
            const double pr = child2 ? resP : 0.0;
            const double2 o2 = first ? make_double2(res, pr) : make_double2(pr, res);
            if (C.st == 16) *reinterpret_cast<double2 *>(cbq + (long)csy * ((gv + 1) >> 1)) = o2;
            else cbq[(long)csy * ((gv + 1) >> 1)] = o2.x;
          }
        }
        resP = res;
      }
    }
    issue<(K + 1 + RP_D) % RD>(t + 1 + RP_D);
  }
};

// ---- up --------------------------------------------------------------------------------------
// coarse-grid correction, NSW sweeps (colour 1 then 0, each).
//   NSW = 1: colour 1 of row t, colour 0 of row t-2 (independent of each other), store of row t-2;
//   NSW = 2: colours 1,0,1,0 of rows t, t-1, t-2, t-3, store of row t-3.
template <bool WRAP, int NSW>
struct Up {
  static constexpr int HY = NSW == 1 ? 2 : 4;
  static constexpr int HX = WRAP ? 0 : (NSW == 1 ? 2 : 4);
  static constexpr int TX = RW - 2 * HX, CO = 192;
  static constexpr int DOUT = NSW == 1 ? 2 : 3;
  static constexpr int RPU = RDU - DOUT - 1;   // rows in flight (a slot must outlive step r + DOUT): 5 / 4 for RDU = 8
  const RLevel &L;
  ColPair C;
  double *ring;
  unsigned rbase, cbase;
  const double *xsrc, *bsrc, *csrc;
  double *xdst;
  int csy, lane, j0, mry, HT, c1, c2;
  double xa[RDU], xb[RDU];

  __device__ __forceinline__ Up(const RLevel &L_) : L(L_) {}

  template <int SL>
  __device__ __forceinline__ void issue(int t) {
    const int gj = ring_row<HY>(L, j0 - HY + t);
    const bool on = gj && t < HT;
    const int nb = on ? C.ld : 0;
    const long o = (long)L.sy * gj;
    cp_async16(rbase + SL * SLOT_U * 8, xsrc + o, nb);
    cp_async16(rbase + (SL * SLOT_U + 64) * 8, bsrc + o, nb);
    const long co = (long)csy * (c2 ? (gj + 1) >> 1 : gj);
    if (c1) cp_async8(rbase + (SL * SLOT_U + 128) * 8, csrc + co, nb ? 8 : 0);
    else cp_async16(rbase + (SL * SLOT_U + 128) * 8, csrc + co, nb);
    if (lane < 4) cp_async16(cbase + SL * SLOT_U * 8, L.rtab + 8 * (on ? gj : 0) + 2 * lane, 16);
    cp_commit();
  }

  template <int K, int D, int COLOR>
  __device__ __forceinline__ void upd() {
    constexpr int R = (K + RDU - D) % RDU, Rm = (K + 2 * RDU - D - 1) % RDU, Rp = (K + RDU - D + 1) % RDU;
    constexpr bool first0 = ((K + RDU - D + HY) & 1) == 0;
    constexpr bool A = (COLOR == 0) == first0;
    const double *slot = ring + R * SLOT_U;
    if (A) xa[R] = ring_upd<WRAP, CO>(true, C, slot, lane, xb[R], xa[Rm], xa[Rp]);
    else xb[R] = ring_upd<WRAP, CO>(false, C, slot, lane, xa[R], xb[Rm], xb[Rp]);
  }

  template <int K>
  __device__ __forceinline__ void step(int t0) {
    const int t = t0 + K;
    {
      constexpr int S = (K + 1) % RDU;
      cp_wait<RPU - 1>();
      __syncwarp();
      const double *slot = ring + S * SLOT_U;
      const double2 v = *reinterpret_cast<const double2 *>(slot + 2 * lane);
      const double2 c = *reinterpret_cast<const double2 *>(slot + 128 + 2 * lane);
      xa[S] = v.x + c.x;
      xb[S] = C.ld == 16 ? v.y + (c1 ? c.x : c.y) : 0.0;
    }
    if (NSW == 1) {
      upd<K, 0, 1>();
      upd<K, 2, 0>();
    } else {
      upd<K, 0, 1>();
      upd<K, 1, 0>();
      upd<K, 2, 1>();
      upd<K, 3, 0>();
    }
    {
      constexpr int R = (K + RDU - DOUT) % RDU;
      const int gv = j0 + t - DOUT - HY;
      if (t >= HY + DOUT && t < HY + DOUT + mry && gv <= L.n2) {
        const long o = (long)L.sy * gv;
        if (C.st == 16) *reinterpret_cast<double2 *>(xdst + o) = make_double2(xa[R], xb[R]);
        else if (C.st == 8) xdst[o] = xa[R];
      }
    }
    issue<(K + 1 + RPU) % RDU>(t + 1 + RPU);
  }
};

}  // namespace

// HASX: the incoming iterate exists (level 0); below it the iterate is zero on entry
template <bool WRAP, bool HASX, int CM, int NSW>
__global__ void __launch_bounds__(WARPS_D * 32, SNAC_RING_MINB_D) mg_ring_down_kernel(const RLevel L, const double *__restrict__ xin, double *__restrict__ xout,
                                                                    const double *__restrict__ b, double *__restrict__ cb, int csy, long csz,
                                                                    const double *lamq, double alpha, double *norms, int nq, int want_b,
                                                                    int ntx, int nty, int mry, const int *active, double xscale) {
  using D = Down<WRAP, HASX, CM, NSW>;
  extern __shared__ __align__(16) double ring_all[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long wid = (long)blockIdx.x * WARPS_D + wib;
  const int tx = (int)(wid % ntx), ty = (int)((wid / ntx) % nty), q = (int)(wid / ((long)ntx * nty));
  if (q >= nq || (active && !active[q])) return;
  D s(L);
  s.lane = lane;
  s.j0 = 1 + ty * mry;
  s.mry = mry;
  s.HT = mry + 2 * D::HY;
  s.want_b = want_b;
  s.csy = csy;
  s.xscale = xscale;
  const long qoff = (long)(q + 1) * L.sz;
  ring_cols<D::HX, WRAP>(L, 1 + tx * D::TX, alpha + __ldg(lamq + q + 1), lane, s.C);
  s.ring = ring_all + (size_t)wib * RD * SLOT_D;
  s.rbase = (unsigned)__cvta_generic_to_shared(s.ring + 2 * lane);
  s.cbase = (unsigned)__cvta_generic_to_shared(s.ring + D::CO + 2 * (lane & 3));
  s.xsrc = (HASX ? xin : b) + qoff + s.C.gi0;
  s.bsrc = b + qoff + s.C.gi0;
  s.xdst = xout + qoff + s.C.gi0;
  s.cbq = cb ? cb + (long)(q + 1) * csz + (CM == CM_Y ? s.C.gi0 : (s.C.gi0 + 1) >> 1) : nullptr;
  s.wc0 = s.C.st ? s.C.w0 : 0.0;        // residual weights: 0 for columns not kept
  s.wc1 = s.C.st == 16 ? s.C.w1 : 0.0;
#pragma unroll
  for (int i = 0; i < RD; ++i) s.xa[i] = s.xb[i] = 0.0;
  // the ring slots of rows "before the first" are read by the first steps: they must hold finite numbers
  for (int i = lane; i < RD * SLOT_D; i += 32) s.ring[i] = 0.0;
  __syncwarp();
  // rows 0 .. RP_D are in flight before the march starts
  s.template issue<0>(0); s.template issue<1>(1); s.template issue<2>(2);
  if (RP_D >= 3) s.template issue<3>(3);
  if (RP_D >= 4) s.template issue<4>(4);
  if (RP_D >= 5) s.template issue<5>(5);
  if (RP_D >= 6) s.template issue<6>(6);
  {  // row 0 enters the window
    cp_wait<RP_D - 1>();
    __syncwarp();
    if (HASX) {
      const double2 v = *reinterpret_cast<const double2 *>(s.ring + 2 * lane);
      s.xa[0] = v.x * xscale; s.xb[0] = v.y * xscale;
    }
  }
  for (int t0 = 0; t0 < s.mry + D::HY + D::DOUT; t0 += RD) {   // the last kept row leaves at step HY + mry - 1 + DOUT
    s.template step<0>(t0); s.template step<1>(t0); s.template step<2>(t0); s.template step<3>(t0);
    s.template step<4>(t0); s.template step<5>(t0); s.template step<6>(t0); s.template step<7>(t0);
    if (RD >= 10) { s.template step<8 % RD>(t0); s.template step<9 % RD>(t0); }
    if (RD == 12) { s.template step<10 % RD>(t0); s.template step<11 % RD>(t0); }
  }
  cp_wait<0>();
  if (norms) {
    double r2 = s.r2, b2 = s.b2;
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

template <bool WRAP, int NSW>
__global__ void __launch_bounds__(WARPS_U * 32, SNAC_RING_MINB_U) mg_ring_up_kernel(const RLevel L, const double *__restrict__ xin, double *__restrict__ xout,
                                                                  const double *__restrict__ b, const double *__restrict__ cx, int csy,
                                                                  long csz, const double *lamq, double alpha, int nq, int ntx, int nty,
                                                                  int mry, const int *active) {
  using U = Up<WRAP, NSW>;
  extern __shared__ __align__(16) double ring_all[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long wid = (long)blockIdx.x * WARPS_U + wib;
  const int tx = (int)(wid % ntx), ty = (int)((wid / ntx) % nty), q = (int)(wid / ((long)ntx * nty));
  if (q >= nq || (active && !active[q])) return;
  U s(L);
  s.lane = lane;
  s.j0 = 1 + ty * mry;
  s.mry = mry;
  s.HT = mry + 2 * U::HY;
  s.csy = csy;
  s.c1 = L.c1; s.c2 = L.c2;
  const long qoff = (long)(q + 1) * L.sz;
  ring_cols<U::HX, WRAP>(L, 1 + tx * U::TX, alpha + __ldg(lamq + q + 1), lane, s.C);
  s.ring = ring_all + (size_t)wib * RDU * SLOT_U;
  s.rbase = (unsigned)__cvta_generic_to_shared(s.ring + 2 * lane);
  s.cbase = (unsigned)__cvta_generic_to_shared(s.ring + U::CO + 2 * (lane & 3));
  s.xsrc = xin + qoff + s.C.gi0;
  s.bsrc = b + qoff + s.C.gi0;
  s.csrc = cx + (long)(q + 1) * csz + (L.c1 ? (s.C.gi0 + 1) >> 1 : s.C.gi0);
  s.xdst = xout + qoff + s.C.gi0;
#pragma unroll
  for (int i = 0; i < RDU; ++i) s.xa[i] = s.xb[i] = 0.0;
  for (int i = lane; i < RDU * SLOT_U; i += 32) s.ring[i] = 0.0;
  __syncwarp();
  static_assert(U::RPU >= 2 && U::RPU <= 5, "prologue below");
  s.template issue<0>(0); s.template issue<1>(1); s.template issue<2>(2);
  if (U::RPU >= 3) s.template issue<3>(3);
  if (U::RPU >= 4) s.template issue<4>(4);
  if (U::RPU >= 5) s.template issue<5>(5);
  {
    cp_wait<U::RPU - 1>();
    __syncwarp();
    const double2 v = *reinterpret_cast<const double2 *>(s.ring + 2 * lane);
    const double2 c = *reinterpret_cast<const double2 *>(s.ring + 128 + 2 * lane);
    s.xa[0] = v.x + c.x;
    s.xb[0] = s.C.ld == 16 ? v.y + (s.c1 ? c.x : c.y) : 0.0;
  }
  for (int t0 = 0; t0 < s.mry + U::HY + U::DOUT; t0 += RDU) {   // the last kept row leaves at step HY + mry - 1 + DOUT
    s.template step<0>(t0); s.template step<1>(t0); s.template step<2>(t0); s.template step<3>(t0);
    s.template step<4>(t0); s.template step<5>(t0);
    if (RDU == 8) { s.template step<6 % RDU>(t0); s.template step<7 % RDU>(t0); }
  }
  cp_wait<0>();
}

// rows per chunk.  Every warp of a launch marches the same number of steps at a pace set by latencies,
// not by throughput, so the launch lasts (waves of CTAs) x (steps per chunk): the chunk height is the
// one that minimises that product for the planes still iterating (nact: an estimate), given the
// CTA slots of the device (slots = SMs x resident CTAs per SM of this kernel).  Short chunks recompute
// relatively more halo rows, long ones leave the last wave half empty.  The chunk height plus the
// steps spent on its halos and on draining the pipeline is a multiple of the unroll period.
static int ring_rows(int n1, int n2, int nact, int tx, int fixed, int period, int warps_per_cta, long slots) {
  const int ntx = (n1 + tx - 1) / tx;
  fixed = (fixed + 1) & ~1;            // steps of a chunk = mry + fixed: a multiple of the (even) period, mry even
  const int smin = 2 * period, smax = 264 / period * period;
  double best = 1e300;
  int best_mry = smin - fixed;
  for (int st = smin; st <= smax; st += period) {
    int mry = st - fixed;
    if (mry < 2) continue;
    const int nty = (n2 + mry - 1) / mry;
    // the shortest chunk height that still needs only nty chunks (balanced chunks)
    while (mry - period >= 2 && (n2 + (mry - period) - 1) / (mry - period) == nty) mry -= period;
    const long ctas = ((long)ntx * nty * std::max(nact, 1) + warps_per_cta - 1) / warps_per_cta;
    const long waves = (ctas + slots - 1) / slots;
    // a last wave that is nearly empty costs as much as a full one; add a small share of the total
    // work so that, between equal wave counts, the launch with less recomputation wins
    const int steps = mry + fixed + 6;   // + set-up of a warp, in units of a step
    const double cost = (double)waves * steps + 0.05 * (double)ctas / slots * steps;
    if (cost < best) { best = cost; best_mry = mry; }
    if (nty == 1) break;
  }
  return best_mry;
}

template <class K>
static long ring_slots(K kernel, int threads, size_t smem) {
  int dev = 0, sms = 148, per_sm = 1;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
  return (long)sms * per_sm;
}

// opt in to more than 48 KB of dynamic shared memory, once per kernel (TAG makes the flag per kernel)
template <int TAG, class K>
static void ring_smem(K kernel, size_t bytes) {
  static bool done = false;
  if (!done) {
    SNAC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    done = true;
  }
}

template <bool WRAP, bool HASX, int CM, int NSW>
static void ring_down_go(cudaStream_t st, const RLevel &L, const double *xin, double *xout, const double *b, double *cb, int csy,
                         long csz, const double *lamq, double alpha, double *norms, int nq, int want_b, const int *active, int nact,
                         double xscale) {
  using D = Down<WRAP, HASX, CM, NSW>;
  constexpr size_t smem = (size_t)WARPS_D * RD * SLOT_D * sizeof(double);
  ring_smem<WRAP * 16 + HASX * 8 + CM * 2 + (NSW - 1)>(mg_ring_down_kernel<WRAP, HASX, CM, NSW>, smem);
  static const long slots = ring_slots(mg_ring_down_kernel<WRAP, HASX, CM, NSW>, WARPS_D * 32, smem);
  const int mry = ring_rows(L.n1, L.n2, nact, D::TX, D::HY + D::DOUT, RD, WARPS_D, slots);
  const int ntx = (L.n1 + D::TX - 1) / D::TX, nty = (L.n2 + mry - 1) / mry;
  const long nw = (long)ntx * nty * nq;
  const unsigned grid = (unsigned)((nw + WARPS_D - 1) / WARPS_D);
  mg_ring_down_kernel<WRAP, HASX, CM, NSW><<<grid, WARPS_D * 32, smem, st>>>(L, xin, xout, b, cb, csy, csz, lamq, alpha, norms, nq, want_b,
                                                                             ntx, nty, mry, active, xscale);
}

#define RING_DOWN_ARGS st, L, xin, xout, b, cb, (int)csy, csz, lamq, alpha, norms, nq, want_b, active, nact, xscale
template <bool WRAP, bool HASX, int NSW>
static void ring_down_cm(int cm, cudaStream_t st, const RLevel &L, const double *xin, double *xout, const double *b, double *cb, long csy,
                         long csz, const double *lamq, double alpha, double *norms, int nq, int want_b, const int *active, int nact,
                         double xscale) {
  if (cm == CM_FULL) ring_down_go<WRAP, HASX, CM_FULL, NSW>(RING_DOWN_ARGS);
  else if (cm == CM_X) ring_down_go<WRAP, HASX, CM_X, NSW>(RING_DOWN_ARGS);
  else ring_down_go<WRAP, HASX, CM_Y, NSW>(RING_DOWN_ARGS);
}

template <int NSW>
static void ring_down_nsw(cudaStream_t st, const RLevel &L, const double *xin, double *xout, const double *b, double *cb, long csy,
                          long csz, const double *lamq, double alpha, double *norms, int nq, int want_b, const int *active, int nact,
                          double xscale) {
  const bool wrap = L.per1 && L.n1 == RW;
  const int cm = (L.c1 && L.c2) ? CM_FULL : (L.c1 ? CM_X : CM_Y);
  if (wrap) {
    if (xin) ring_down_cm<true, true, NSW>(cm, RING_DOWN_ARGS);
    else ring_down_cm<true, false, NSW>(cm, RING_DOWN_ARGS);
  } else {
    if (xin) ring_down_cm<false, true, NSW>(cm, RING_DOWN_ARGS);
    else ring_down_cm<false, false, NSW>(cm, RING_DOWN_ARGS);
  }
}

void launch_ring_down(cudaStream_t st, const RLevel &L, int nsweep, const double *xin, double *xout, const double *b, double *cb, long csy,
                      long csz, const double *lamq, double alpha, double *norms, int nq, int want_b, const int *active, int nact,
                      double xscale) {
  if (nsweep == 1) ring_down_nsw<1>(RING_DOWN_ARGS);
  else ring_down_nsw<2>(RING_DOWN_ARGS);
}
#undef RING_DOWN_ARGS

template <bool WRAP, int NSW>
static void ring_up_go(cudaStream_t st, const RLevel &L, const double *xin, double *xout, const double *b, const double *cx, long csy,
                       long csz, const double *lamq, double alpha, int nq, const int *active, int nact) {
  using U = Up<WRAP, NSW>;
  constexpr size_t smem = (size_t)WARPS_U * RDU * SLOT_U * sizeof(double);
  ring_smem<100 + WRAP * 2 + (NSW - 1)>(mg_ring_up_kernel<WRAP, NSW>, smem);
  static const long slots = ring_slots(mg_ring_up_kernel<WRAP, NSW>, WARPS_U * 32, smem);
  const int mry = ring_rows(L.n1, L.n2, nact, U::TX, U::HY + U::DOUT, RDU, WARPS_U, slots);
  const int ntx = (L.n1 + U::TX - 1) / U::TX, nty = (L.n2 + mry - 1) / mry;
  const long nw = (long)ntx * nty * nq;
  const unsigned grid = (unsigned)((nw + WARPS_U - 1) / WARPS_U);
  mg_ring_up_kernel<WRAP, NSW><<<grid, WARPS_U * 32, smem, st>>>(L, xin, xout, b, cx, (int)csy, csz, lamq, alpha, nq, ntx, nty, mry, active);
}

void launch_ring_up(cudaStream_t st, const RLevel &L, int nsweep, const double *xin, double *xout, const double *b, const double *cx,
                    long csy, long csz, const double *lamq, double alpha, int nq, const int *active, int nact) {
  const bool wrap = L.per1 && L.n1 == RW;
  if (nsweep == 1) {
    if (wrap) ring_up_go<true, 1>(st, L, xin, xout, b, cx, csy, csz, lamq, alpha, nq, active, nact);
    else ring_up_go<false, 1>(st, L, xin, xout, b, cx, csy, csz, lamq, alpha, nq, active, nact);
  } else {
    if (wrap) ring_up_go<true, 2>(st, L, xin, xout, b, cx, csy, csz, lamq, alpha, nq, active, nact);
    else ring_up_go<false, 2>(st, L, xin, xout, b, cx, csy, csz, lamq, alpha, nq, active, nact);
  }
}

}  // namespace snac
