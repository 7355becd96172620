// K3a and K6, z-marching form: the Poisson right-hand side (fillps + updt_rhs) and the projection + pressure update
// (correc + updt_pressure) with their planes staged in shared memory by TMA.
//
// Same arithmetic as fillps_kernel / correc_kernel (pois.cu; reference src/fillps.f90:29-53, src/bound.f90:543-585,
// src/correc.f90:23-47, src/updt_pressure.f90:23-46), data movement of rk_mom_tma_kernel (mom_tma.cu): a CTA owns an
// x-y tile of 32 x 8 cells and marches over a chunk of z-planes; one thread issues, NS-2 planes ahead of their use,
// one cp.async.bulk.tensor box load per field and plane into a ring of shared-memory slots (one mbarrier per slot).
// The one-thread-per-cell kernels were bound by the latency of their global loads (long-scoreboard stalls at 88 %
// occupancy: 0.62 and 0.75 of the measured HBM copy rate); here a thread only reads shared memory and writes its
// results with aligned, coalesced stores, and what it needs from the neighbouring plane along z is handed over in
// registers.
#include "tma.cuh"

namespace snac {

namespace {

using tma::mbar_expect_tx;
using tma::mbar_init;
using tma::mbar_wait;

constexpr int TX = 32, TY = 8;           // cells per tile = threads per CTA

// ---- K3a ------------------------------------------------------------------------------------------------------
// per slot: u tile with the column i0-1 (box from i0-2: tma.cuh), v tile with the row j0-1, w tile
constexpr int FUX = TX + 2;
constexpr int FU = FUX * TY, FV = TX * (TY + 1), FW = TX * TY;
constexpr int FSLOT = FU + FV + FW;      // doubles; every part is a multiple of 128 bytes
static_assert((FU * 8) % 128 == 0 && (FV * 8) % 128 == 0 && (FW * 8) % 128 == 0, "TMA destinations are 128-byte aligned");
#ifndef SNAC_POIS_KZ
#define SNAC_POIS_KZ 16
#endif
#ifndef SNAC_FILL_NS
#define SNAC_FILL_NS 10
#endif
constexpr int FNS = SNAC_FILL_NS;

struct FillTmaArgs {
  int n[3];
  long sy, sz;
  double *rhs;
  const double *wq[3], *gq[3];
  double cb[3][2];
  double rdt;
  int do_bc;
  int kz;              // planes per chunk
};

__global__ void __launch_bounds__(TX * TY, 3) fillps_tma_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
                                                                const __grid_constant__ CUtensorMap tm_w, const FillTmaArgs a) {
  extern __shared__ __align__(128) double ring[];   // [FNS][u | v | w]
  __shared__ unsigned long long full[FNS];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int k0 = 1 + blockIdx.z * a.kz, k1 = min(k0 + a.kz - 1, a.n[2]);
  const int i = i0 + tx, j = j0 + ty;
  const bool in = i <= a.n[0] && j <= a.n[1];
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < FNS; ++s) mbar_init(&full[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
  // plane p (numbered from k0-1 = 0) lives in slot p mod FNS; plane 0 only hands w(k0-1) over
  const int np = k1 - k0 + 2;
  auto issue = [&](int p) {
    const int s = p % FNS;
    double *dst = ring + (size_t)s * FSLOT;
    const int kk = k0 - 1 + p;
    mbar_expect_tx(&full[s], (p > 0 ? (unsigned)FSLOT : (unsigned)FW) * 8u);
    if (p > 0) {
      tma::load_3d(dst, &tm_u, &full[s], i0 - 1, j0, kk);            // ghosted map: element i0-2
      tma::load_3d(dst + FU, &tm_v, &full[s], i0 - 1, j0 - 1, kk);   // interior maps: element i0
    }
    tma::load_3d(dst + FU + FV, &tm_w, &full[s], i0 - 1, j0, kk);
  };
  if (tid == 0)
    for (int p = 0; p < min(FNS, np); ++p) issue(p);

  const int ic = min(i, a.n[0]), jc = min(j, a.n[1]);
  const double wx = __ldg(a.wq[0] + ic), wy = __ldg(a.wq[1] + jc);
  const double gx = __ldg(a.gq[0] + ic), gy = __ldg(a.gq[1] + jc);
  // updt_rhs on the faces normal to x and y (src/bound.f90:543-585); the weight along z is applied per plane
  double bxy = 0.0;
  if (a.do_bc) {
    if (i == 1) bxy += a.cb[0][0] * wy;
    if (i == a.n[0]) bxy += a.cb[0][1] * wy;
    if (j == 1) bxy += a.cb[1][0] * wx;
    if (j == a.n[1]) bxy += a.cb[1][1] * wx;
  }
  const int ou = ty * FUX + tx + 2, ov = FU + (ty + 1) * TX + tx, ow = FU + FV + ty * TX + tx;
  const long c0 = ic + a.sy * jc;
  mbar_wait(&full[0], 0);
  double wkm = ring[ow];
  for (int k = k0; k <= k1; ++k) {
    const int p = k - k0 + 1;
    __syncthreads();   // every thread is done with plane p-1: its slot takes plane p-1+FNS
    if (tid == 0 && p - 1 + FNS < np) issue(p - 1 + FNS);
    const double *s = ring + (size_t)(p % FNS) * FSLOT;
    mbar_wait(&full[p % FNS], (p / FNS) & 1);
    const double wz = __ldg(a.wq[2] + k), gz = __ldg(a.gq[2] + k);
    const double du = s[ou] - s[ou - 1], dv = s[ov] - s[ov - TX], w0 = s[ow], dw = w0 - wkm;
    wkm = w0;
    double val = (dw * gz * wx * wy + dv * gy * wx * wz + du * gx * wy * wz) * a.rdt;
    if (a.do_bc) {
      val += bxy * wz;
      if (k == 1) val += a.cb[2][0] * wx * wy;
      if (k == a.n[2]) val += a.cb[2][1] * wx * wy;
    }
    if (in) a.rhs[c0 + a.sz * k] = val;
  }
}

// ---- K6 -------------------------------------------------------------------------------------------------------
// per slot: pp tile with halo (box from (i0-2, j0-1), as the tiles of rk_mom_tma_kernel), interior tiles of u, v, w, p
constexpr int BX = TX + 4, BY = TY + 2;
constexpr int CPP = ((BX * BY * 8 + 127) / 128) * 128 / 8, CF = TX * TY;
constexpr int CSLOT = CPP + 4 * CF;
#ifndef SNAC_CORR_NS
#define SNAC_CORR_NS 8
#endif
constexpr int CNS = SNAC_CORR_NS;

struct CorrTmaArgs {
  int n[3];
  long sy, sz;
  double *u, *v, *w, *p;
  const double *rdc[3], *rdf[3];
  double dtrk, alpha;
  int do_correc, do_updt;
  int kz;
};

// The ghost columns i = 0 and (when n is a multiple of the tile width) i = n+1 lie outside the 128-byte aligned interior
// tiles; the reference's ranges include them (src/correc.f90:23-47).  They are corrected by a kernel of their own, one
// thread per ghost cell: done by a lane of the first / last tile inside the z-march, their global read-modify-writes
// sat in the critical path of every plane step of those CTAs (a quarter of all CTAs on a 256^3 box: 4.4 instead of
// 6 TB/s).  Same arithmetic; pp is only read, so the order of the two kernels does not matter.
__global__ void __launch_bounds__(128) correc_ghost_columns_kernel(const CorrTmaArgs a, const double *pp, int both) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. n1+1
  const int k = blockIdx.y;                              // 0 .. n2+1
  if (j > a.n[1] + 1) return;
  for (int side = 0; side <= both; ++side) {
    const int i = side == 0 ? 0 : a.n[0] + 1;
    const long c = i + a.sy * j + a.sz * k;
    const double p0 = pp[c];
    if (i <= a.n[0]) a.u[c] -= a.dtrk * (pp[c + 1] - p0) * __ldg(a.rdc[0] + i);
    if (j <= a.n[1]) a.v[c] -= a.dtrk * (pp[c + a.sy] - p0) * __ldg(a.rdc[1] + j);
    if (k <= a.n[2]) a.w[c] -= a.dtrk * (pp[c + a.sz] - p0) * __ldg(a.rdc[2] + k);
  }
}

template <bool IMPDIFF>
__global__ void __launch_bounds__(TX * TY, 2) correc_tma_kernel(const __grid_constant__ CUtensorMap tm_pp, const __grid_constant__ CUtensorMap tm_u,
                                                                const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_w,
                                                                const __grid_constant__ CUtensorMap tm_p, const CorrTmaArgs a) {
  extern __shared__ __align__(128) double ring[];   // [CNS][pp with halo | u | v | w | p]
  __shared__ unsigned long long full[CNS];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int n0 = a.n[0], n1 = a.n[1], n2 = a.n[2];
  const int i0 = 1 + blockIdx.x * TX, j0 = blockIdx.y * TY;          // rows from the ghost row 0, planes from the ghost plane 0
  const int k0 = blockIdx.z * a.kz, k1 = min(k0 + a.kz - 1, n2 + 1);
  const int i = i0 + tx, j = j0 + ty;
  const bool valid = i <= n0 + 1 && j <= n1 + 1;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < CNS; ++s) mbar_init(&full[s], 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
  // planes k0-1 .. k1+1 of pp, k0 .. k1 of the fields; plane p (numbered from k0-1 = 0) lives in slot p mod CNS
  const int np = k1 - k0 + 3;
  auto issue = [&](int p) {
    const int s = p % CNS;
    double *dst = ring + (size_t)s * CSLOT;
    const int kk = k0 - 1 + p;
    const bool fields = p >= 1 && p <= np - 2;
    mbar_expect_tx(&full[s], (unsigned)(BX * BY + (fields ? 4 * CF : 0)) * 8u);
    tma::load_3d(dst, &tm_pp, &full[s], i0 - 1, j0 - 1, kk);
    if (fields) {
      tma::load_3d(dst + CPP, &tm_u, &full[s], i0 - 1, j0, kk);
      tma::load_3d(dst + CPP + CF, &tm_v, &full[s], i0 - 1, j0, kk);
      tma::load_3d(dst + CPP + 2 * CF, &tm_w, &full[s], i0 - 1, j0, kk);
      tma::load_3d(dst + CPP + 3 * CF, &tm_p, &full[s], i0 - 1, j0, kk);
    }
  };
  if (tid == 0)
    for (int p = 0; p < min(CNS, np); ++p) issue(p);

  const int ic = min(i, n0 + 1), jc = min(j, n1 + 1);
  const double rdxc = __ldg(a.rdc[0] + ic), rdyc = __ldg(a.rdc[1] + jc);
  double rdxc_m = 0.0, rdyc_m = 0.0, rdxf = 0.0, rdyf = 0.0;
  if (IMPDIFF) {
    rdxc_m = __ldg(a.rdc[0] + max(ic - 1, 0)); rdyc_m = __ldg(a.rdc[1] + max(jc - 1, 0));
    rdxf = __ldg(a.rdf[0] + ic); rdyf = __ldg(a.rdf[1] + jc);
  }
  const int o = (ty + 1) * BX + tx + 2, of = CPP + ty * TX + tx;
  const long c0 = ic + a.sy * jc;
  const bool inxy = i >= 1 && i <= n0 && j >= 1 && j <= n1;

  mbar_wait(&full[0], 0);
  double pkm = ring[o];
  mbar_wait(&full[1 % CNS], 0);
  double p0 = ring[(size_t)(1 % CNS) * CSLOT + o];

  for (int k = k0; k <= k1; ++k) {
    const int p = k - k0 + 1;
    __syncthreads();   // every thread is done with plane p-1: its slot takes plane p-1+CNS
    if (tid == 0 && p - 1 + CNS < np) issue(p - 1 + CNS);
    const double *sc = ring + (size_t)(p % CNS) * CSLOT;
    const double *sn = ring + (size_t)((p + 1) % CNS) * CSLOT;
    mbar_wait(&full[(p + 1) % CNS], ((p + 1) / CNS) & 1);   // plane p was waited for one step earlier (or above)
    const double pip = sc[o + 1], pjp = sc[o + BX], pkp = sn[o];
    const long c = c0 + a.sz * k;
    if (valid) {
      if (a.do_correc) {
        if (i <= n0) a.u[c] = sc[of] - a.dtrk * (pip - p0) * rdxc;
        if (j <= n1) a.v[c] = sc[of + CF] - a.dtrk * (pjp - p0) * rdyc;
        if (k <= n2) a.w[c] = sc[of + 2 * CF] - a.dtrk * (pkp - p0) * __ldg(a.rdc[2] + k);
      }
      if (a.do_updt && inxy && k >= 1 && k <= n2) {
        double val = sc[of + 3 * CF] + p0;
        if (IMPDIFF) {   // src/updt_pressure.f90:30-46
          const double pim = sc[o - 1], pjm = sc[o - BX];
          val += a.alpha * (((pip - p0) * rdxc - (p0 - pim) * rdxc_m) * rdxf + ((pjp - p0) * rdyc - (p0 - pjm) * rdyc_m) * rdyf +
                            ((pkp - p0) * __ldg(a.rdc[2] + k) - (p0 - pkm) * __ldg(a.rdc[2] + k - 1)) * __ldg(a.rdf[2] + k));
        }
        a.p[c] = val;
      }
    }
    pkm = p0; p0 = pkp;
  }
}

// chunks along z: short, so that the CTAs of a wave stay in step and find the halos of their tiles in L2 (mom_tma.cu)
// (shorter ones, down to 8 planes, when a small box would not fill the device otherwise)
int planes_per_chunk(int tiles, int nplanes, int &nchunk) {
  nchunk = std::max(1, std::max(nplanes / SNAC_POIS_KZ, std::min(nplanes / 8, (444 + tiles - 1) / tiles)));
  const int kz = (nplanes + nchunk - 1) / nchunk;
  nchunk = (nplanes + kz - 1) / kz;
  return kz;
}

}  // namespace

// tensor maps of the box's fields for the two kernels; false: the one-thread-per-cell kernels are used
bool pois_tma_setup(Box &b) {
  bool ok = true;
  for (int q = 0; q < 2 && ok; ++q) {
    ok = tma_make_map(b.tmap_fill[q][0], b.u[q].p, b.g, FUX, TY, false) && tma_make_map(b.tmap_fill[q][1], b.v[q].p, b.g, TX, TY + 1, true) &&
         tma_make_map(b.tmap_int[q][0], b.u[q].p, b.g, TX, TY, true) && tma_make_map(b.tmap_int[q][1], b.v[q].p, b.g, TX, TY, true) &&
         tma_make_map(b.tmap_int[q][2], b.w[q].p, b.g, TX, TY, true);
  }
  ok = ok && tma_make_map(b.tmap_int[0][3], b.p.p, b.g, TX, TY, true) && tma_make_map(b.tmap_pp, b.pp_ptr(), b.g, BX, BY, false);
  return ok;
}

bool launch_fillps_tma(Ctx &ctx, Box &b, double dtrk, bool do_bc) {
  static const bool off = getenv("SNAC_POIS_TMA") && atoi(getenv("SNAC_POIS_TMA")) == 0;   // A/B switch against pois.cu
  if (off || !b.tma_pois_ok) return false;
  FillTmaArgs a;
  for (int d = 0; d < 3; ++d) {
    a.n[d] = b.g.n[d];
    a.wq[d] = b.wq[d].p(); a.gq[d] = b.gq[d].p();
    a.cb[d][0] = b.cbp[d][0]; a.cb[d][1] = b.cbp[d][1];
  }
  a.sy = b.g.sy; a.sz = b.g.sz;
  a.rhs = b.rhs.p;
  a.rdt = 1.0 / dtrk;
  a.do_bc = do_bc;
  const int ntx = (b.g.n[0] + TX - 1) / TX, nty = (b.g.n[1] + TY - 1) / TY;
  int nchunk;
  a.kz = planes_per_chunk(ntx * nty, b.g.n[2], nchunk);
  const size_t smem = (size_t)FNS * FSLOT * sizeof(double);
  static bool attr = false;
  if (!attr) {
    SNAC_CUDA(cudaFuncSetAttribute(fillps_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  const int q = ctx.par;
  fillps_tma_kernel<<<dim3(ntx, nty, nchunk), dim3(TX, TY, 1), smem, ctx.stream>>>(
      *reinterpret_cast<const CUtensorMap *>(b.tmap_fill[q][0]), *reinterpret_cast<const CUtensorMap *>(b.tmap_fill[q][1]),
      *reinterpret_cast<const CUtensorMap *>(b.tmap_int[q][2]), a);
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
  return true;
}

bool launch_correc_tma(Ctx &ctx, Box &b, double dtrk, double alpha, bool do_correc, bool do_updt) {
  static const bool off = getenv("SNAC_POIS_TMA") && atoi(getenv("SNAC_POIS_TMA")) == 0;
  if (off || !b.tma_pois_ok) return false;
  CorrTmaArgs a;
  for (int d = 0; d < 3; ++d) {
    a.n[d] = b.g.n[d];
    a.rdc[d] = b.rdc[d].p(); a.rdf[d] = b.rdf[d].p();
  }
  a.sy = b.g.sy; a.sz = b.g.sz;
  const int q = ctx.par;
  a.u = b.u[q].p; a.v = b.v[q].p; a.w = b.w[q].p; a.p = b.p.p;
  a.dtrk = dtrk; a.alpha = alpha;
  a.do_correc = do_correc; a.do_updt = do_updt;
  const int ntx = (b.g.n[0] + TX - 1) / TX, nty = (b.g.n[1] + 2 + TY - 1) / TY;
  int nchunk;
  a.kz = planes_per_chunk(ntx * nty, b.g.n[2] + 2, nchunk);
  const size_t smem = (size_t)CNS * CSLOT * sizeof(double);
  const CUtensorMap *tpp = reinterpret_cast<const CUtensorMap *>(b.tmap_pp), *tu = reinterpret_cast<const CUtensorMap *>(b.tmap_int[q][0]);
  const CUtensorMap *tv = reinterpret_cast<const CUtensorMap *>(b.tmap_int[q][1]), *tw = reinterpret_cast<const CUtensorMap *>(b.tmap_int[q][2]);
  const CUtensorMap *tp = reinterpret_cast<const CUtensorMap *>(b.tmap_int[0][3]);
  const dim3 grd(ntx, nty, nchunk), blk(TX, TY, 1);
  if (ctx.prm.impdiff) {
    static bool attr = false;
    if (!attr) { SNAC_CUDA(cudaFuncSetAttribute(correc_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    correc_tma_kernel<true><<<grd, blk, smem, ctx.stream>>>(*tpp, *tu, *tv, *tw, *tp, a);
  } else {
    static bool attr = false;
    if (!attr) { SNAC_CUDA(cudaFuncSetAttribute(correc_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    correc_tma_kernel<false><<<grd, blk, smem, ctx.stream>>>(*tpp, *tu, *tv, *tw, *tp, a);
  }
  ctx.launches++;
  if (do_correc) {
    correc_ghost_columns_kernel<<<dim3((b.g.n[1] + 2 + 127) / 128, b.g.n[2] + 2), 128, 0, ctx.stream>>>(a, b.pp_ptr(), (b.g.n[0] % TX) == 0 ? 1 : 0);
    ctx.launches++;
  }
  SNAC_CUDA(cudaGetLastError());
  return true;
}

}  // namespace snac
