// K1, z-marching form: fused momentum right-hand side + low-storage RK3 update with the velocity and pressure
// planes staged in shared memory by TMA.
//
// Same arithmetic as rk_mom_kernel (mom.cu; reference rk_mom src/rk.f90:7-128, mom{x,y,z}_a_vv src/mom.f90:112-219,
// _d :221-327, _p :328-386), different data movement.  A CTA owns an x-y tile of TX x TY cells and marches over
// a chunk of z-planes; a single thread issues, NS-2 planes ahead of their use, one cp.async.bulk.tensor (TMA) box
// load per field and plane -- the tile of u, v, w and p with its halo -- into a ring of NS
// shared-memory slots, completion signalled on one mbarrier per slot.  The stencil reads its in-plane neighbours
// from the slots of planes k and k+1; everything it needs from plane k-1 (u, v, w at the cell, w at i+1 and at
// j+1) it already holds in registers from the previous step, which also hand u0/uim/v0/vjm/w0/p0 from k+1 to k.
// The RK histories and the six outputs are plain coalesced global accesses (the history of plane k+1 is loaded a
// step ahead).  Global traffic is the algorithmic 7 reads + 6 writes per cell plus the halo of the tiles, which
// neighbouring CTAs find in L2; the latency of the 25 stencil loads per cell, which bound the one-thread-per-cell
// kernel (long-scoreboard stalls at 28 warps per SM), is paid once per plane by the TMA unit instead.
#include "tma.cuh"

namespace snac {

namespace {

#ifndef SNAC_MOM_TY
#define SNAC_MOM_TY 8
#endif
#ifndef SNAC_MOM_MINB
#define SNAC_MOM_MINB 2
#endif
constexpr int TX = 32, TY = SNAC_MOM_TY;  // cells per tile = threads per CTA (one cell each per plane)
// tile with halo.  TMA wants the first element of a box row 16-byte aligned in global memory; the arrays put the
// first INTERIOR element of a row on a 128-byte boundary (Grid3), so the halo column i0-1 sits 8 bytes before one:
// the box starts one element further left (i0-2, never used) and is TX+4 wide.
constexpr int BX = TX + 4, BY = TY + 2;
constexpr int BUF = ((BX * BY * 8 + 127) / 128) * 128 / 8;   // doubles per field and slot (128-byte aligned for TMA)
#ifndef SNAC_MOM_NS
#define SNAC_MOM_NS 6
#endif
// planes per chunk of the z-march.  Short chunks keep the CTAs of a wave in step, so that the halo of a tile is still
// in L2 when the neighbouring CTA asks for it: chunks of 256 planes (channel, 1024 x 512 x 512) let them drift apart and
// every halo came from DRAM again (117 B per point, 4.9 TB/s); 27 planes: 5.8 TB/s (profiles/r02_analysis.md)
#ifndef SNAC_MOM_KZ
#define SNAC_MOM_KZ 24
#endif
constexpr int NS = SNAC_MOM_NS;           // ring slots: planes k and k+1 in use, NS - 2 in flight
constexpr unsigned SLOT_BYTES = 4u * BX * BY * 8u;
constexpr int HBUF = TX * TY;             // doubles per RK-history tile (interior only) and slot
constexpr unsigned HIST_BYTES = 3u * HBUF * 8u;
template <bool FIRST>
__host__ __device__ constexpr int slot_doubles() { return 4 * BUF + (FIRST ? 0 : 3 * HBUF); }

struct MomTmaArgs {
  int n[3];
  long sy, sz;
  const double *df[3], *rdc[3], *rdf[3];
  double *un, *vn, *wn, *du, *dv, *dw;
  double f1, f2, f12, visc;
  double bforce[3];
  int kz;              // planes per chunk
};

using tma::mbar_init;
using tma::mbar_expect_tx;
using tma::mbar_wait;
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, unsigned long long *bar, int c0, int c1, int c2) {
  tma::load_3d(dst, map, bar, c0, c1, c2);
}

template <bool IMPDIFF, bool FIRST>
__global__ void __launch_bounds__(TX * TY, SNAC_MOM_MINB) rk_mom_tma_kernel(const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_v,
                                                               const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_p,
                                                               const __grid_constant__ CUtensorMap tm_du, const __grid_constant__ CUtensorMap tm_dv,
                                                               const __grid_constant__ CUtensorMap tm_dw, const MomTmaArgs a) {
  constexpr int SLOT = slot_doubles<FIRST>();
  extern __shared__ __align__(128) double ring[];   // [NS][u, v, w, p tiles with halo | du, dv, dw interior tiles]
  __shared__ unsigned long long full[NS];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
  const int i0 = 1 + blockIdx.x * TX, j0 = 1 + blockIdx.y * TY;
  const int k0 = 1 + blockIdx.z * a.kz;                      // first plane of the chunk
  const int k1 = min(k0 + a.kz - 1, a.n[2]);                 // last one
  const int i = i0 + tx, j = j0 + ty;
  const bool in = i <= a.n[0] && j <= a.n[1];
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
    tma::fence_barrier_init();   // the barriers are visible to the TMA unit
  }
  __syncthreads();
  // planes k0-1 .. k1+1 are needed; plane p (numbered from k0-1 = 0) lives in slot p mod NS
  const int np = k1 - k0 + 3;
  auto issue = [&](int p) {
    const int s = p % NS;
    double *dst = ring + (size_t)s * SLOT;
    const bool hist = !FIRST && p >= 1 && p <= np - 2;        // the RK histories of the planes that are updated
    mbar_expect_tx(&full[s], SLOT_BYTES + (hist ? HIST_BYTES : 0u));
    const int kk = k0 - 1 + p;
    if (hist) {   // interior tiles: coordinate = index - 1 (mom_tma_make_map with interior = true)
      tma_load_3d(dst + 4 * BUF, &tm_du, &full[s], i0 - 1, j0, kk);
      tma_load_3d(dst + 4 * BUF + HBUF, &tm_dv, &full[s], i0 - 1, j0, kk);
      tma_load_3d(dst + 4 * BUF + 2 * HBUF, &tm_dw, &full[s], i0 - 1, j0, kk);
    }
    tma_load_3d(dst, &tm_u, &full[s], i0 - 1, j0 - 1, kk);            // coordinate = index + 1 (see mom_tma_make_map): i0-2
    tma_load_3d(dst + BUF, &tm_v, &full[s], i0 - 1, j0 - 1, kk);
    tma_load_3d(dst + 2 * BUF, &tm_w, &full[s], i0 - 1, j0 - 1, kk);
    tma_load_3d(dst + 3 * BUF, &tm_p, &full[s], i0 - 1, j0 - 1, kk);
  };
  if (tid == 0)
    for (int p = 0; p < min(NS, np); ++p) issue(p);

  // per-thread metrics (src/mom.f90 passes dxc, dxf, ... ; reciprocals precomputed on the host)
  const int ic = min(i, a.n[0]), jc = min(j, a.n[1]);   // clamped: threads outside the box only take part in the barriers
  const double dxf_i = __ldg(a.df[0] + ic), dxf_ip = __ldg(a.df[0] + ic + 1);
  const double dyf_j = __ldg(a.df[1] + jc), dyf_jp = __ldg(a.df[1] + jc + 1);
  const double rdxf_i = __ldg(a.rdf[0] + ic), rdxf_ip = __ldg(a.rdf[0] + ic + 1);
  const double rdyf_j = __ldg(a.rdf[1] + jc), rdyf_jp = __ldg(a.rdf[1] + jc + 1);
  const double rdxc_i = __ldg(a.rdc[0] + ic), rdxc_im = __ldg(a.rdc[0] + ic - 1);
  const double rdyc_j = __ldg(a.rdc[1] + jc), rdyc_jm = __ldg(a.rdc[1] + jc - 1);
  const int o = (ty + 1) * BX + tx + 2;                  // the cell inside a tile
  const long c0 = ic + a.sy * jc;

  // prologue: plane k0-1 (what a step inherits from "k-1") and the centre values of plane k0
  double ukm, vkm, wkm, wipkm, wjpkm, u0, uim, v0, vjm, w0, p0;
  {
    mbar_wait(&full[0], 0);
    const double *s = ring;
    ukm = s[o]; vkm = s[BUF + o]; wkm = s[2 * BUF + o]; wipkm = s[2 * BUF + o + 1]; wjpkm = s[2 * BUF + o + BX];
    mbar_wait(&full[1 % NS], 0);
    const double *t = ring + (size_t)(1 % NS) * SLOT;
    u0 = t[o]; uim = t[o - 1]; v0 = t[BUF + o]; vjm = t[BUF + o - BX]; w0 = t[2 * BUF + o]; p0 = t[3 * BUF + o];
  }
  const int oh = ty * TX + tx;                            // the cell inside a history tile
  double dzf_k = __ldg(a.df[2] + k0), rdzf_k = __ldg(a.rdf[2] + k0), rdzc_km = __ldg(a.rdc[2] + k0 - 1);

  for (int k = k0; k <= k1; ++k) {
    const int p = k - k0 + 1;                            // plane k in chunk numbering; k+1 = p+1
    // every thread is done with plane p-1 (read last during the previous step): its slot takes plane p-1+NS
    __syncthreads();
    if (tid == 0 && p - 1 + NS < np) issue(p - 1 + NS);
    const double *sc = ring + (size_t)(p % NS) * SLOT;
    const double *sn = ring + (size_t)((p + 1) % NS) * SLOT;
    mbar_wait(&full[(p + 1) % NS], ((p + 1) / NS) & 1);   // plane p was waited for one step earlier (or in the prologue)
    const double dzf_kp = __ldg(a.df[2] + k + 1), rdzf_kp = __ldg(a.rdf[2] + k + 1), rdzc_k = __ldg(a.rdc[2] + k);
    const long c = c0 + a.sz * k;
    double hu = 0.0, hv = 0.0, hw = 0.0;
    if (!FIRST) { hu = sc[4 * BUF + oh]; hv = sc[4 * BUF + HBUF + oh]; hw = sc[4 * BUF + 2 * HBUF + oh]; }
    const double uip = sc[o + 1], ujp = sc[o + BX], ujm = sc[o - BX], uimjp = sc[o + BX - 1];
    const double vip = sc[BUF + o + 1], vim = sc[BUF + o - 1], vjp = sc[BUF + o + BX], vipjm = sc[BUF + o + 1 - BX];
    const double wip = sc[2 * BUF + o + 1], wim = sc[2 * BUF + o - 1], wjp = sc[2 * BUF + o + BX], wjm = sc[2 * BUF + o - BX];
    const double pip = sc[3 * BUF + o + 1], pjp = sc[3 * BUF + o + BX];
    const double ukp = sn[o], uimkp = sn[o - 1], vkp = sn[BUF + o], vjmkp = sn[BUF + o - BX], wkp = sn[2 * BUF + o], pkp = sn[3 * BUF + o];

    // advection, volume-weighted form (src/mom.f90:129-142,165-178,201-214)
    const double ax = rdxc_i * 0.25 *
        (((u0 + uim) * uim - (u0 + uip) * uip) +
         rdyf_j * ((vjm * dxf_i + vipjm * dxf_ip) * ujm - (v0 * dxf_i + vip * dxf_ip) * ujp) +
         rdzf_k * ((wkm * dxf_i + wipkm * dxf_ip) * ukm - (w0 * dxf_i + wip * dxf_ip) * ukp));
    const double ay = rdyc_j * 0.25 *
        (rdxf_i * ((uim * dyf_j + uimjp * dyf_jp) * vim - (u0 * dyf_j + ujp * dyf_jp) * vip) +
         ((v0 + vjm) * vjm - (v0 + vjp) * vjp) +
         rdzf_k * ((wkm * dyf_j + wjpkm * dyf_jp) * vkm - (w0 * dyf_j + wjp * dyf_jp) * vkp));
    const double az = rdzc_k * 0.25 *
        (rdxf_i * ((uim * dzf_k + uimkp * dzf_kp) * wim - (u0 * dzf_k + ukp * dzf_kp) * wip) +
         rdyf_j * ((vjm * dzf_k + vjmkp * dzf_kp) * wjm - (v0 * dzf_k + vkp * dzf_kp) * wjp) +
         ((w0 + wkm) * wkm - (w0 + wkp) * wkp));
    // diffusion (src/mom.f90:239-251,275-287,311-323)
    const double dx_ = a.visc * (((uip - u0) * rdxf_ip - (u0 - uim) * rdxf_i) * rdxc_i +
                                 ((ujp - u0) * rdyc_j - (u0 - ujm) * rdyc_jm) * rdyf_j +
                                 ((ukp - u0) * rdzc_k - (u0 - ukm) * rdzc_km) * rdzf_k);
    const double dy_ = a.visc * (((vip - v0) * rdxc_i - (v0 - vim) * rdxc_im) * rdxf_i +
                                 ((vjp - v0) * rdyf_jp - (v0 - vjm) * rdyf_j) * rdyc_j +
                                 ((vkp - v0) * rdzc_k - (v0 - vkm) * rdzc_km) * rdzf_k);
    const double dz_ = a.visc * (((wip - w0) * rdxc_i - (w0 - wim) * rdxc_im) * rdxf_i +
                                 ((wjp - w0) * rdyc_j - (w0 - wjm) * rdyc_jm) * rdyf_j +
                                 ((wkp - w0) * rdzf_kp - (w0 - wkm) * rdzf_k) * rdzc_k);
    // pressure gradient + body force (src/mom.f90:342,361,380)
    const double px = -(pip - p0) * rdxc_i + a.bforce[0];
    const double py = -(pjp - p0) * rdyc_j + a.bforce[1];
    const double pz = -(pkp - p0) * rdzc_k + a.bforce[2];
    double rx, ry, rz;
    if (IMPDIFF) { rx = ax; ry = ay; rz = az; } else { rx = dx_ + ax; ry = dy_ + ay; rz = dz_ + az; }
    double un = u0 + a.f1 * rx, vn = v0 + a.f1 * ry, wn = w0 + a.f1 * rz;
    if (!FIRST) { un += a.f2 * hu; vn += a.f2 * hv; wn += a.f2 * hw; }
    if (IMPDIFF) {  // + f12*d - 0.5*f12*d  (src/rk.f90:75-79,121-123)
      un += a.f12 * dx_; vn += a.f12 * dy_; wn += a.f12 * dz_;
    }
    un += a.f12 * px; vn += a.f12 * py; wn += a.f12 * pz;
    if (IMPDIFF) { un -= 0.5 * a.f12 * dx_; vn -= 0.5 * a.f12 * dy_; wn -= 0.5 * a.f12 * dz_; }
    if (in) {
      a.du[c] = rx; a.dv[c] = ry; a.dw[c] = rz;
      a.un[c] = un; a.vn[c] = vn; a.wn[c] = wn;
    }
    // plane k+1 becomes plane k
    ukm = u0; u0 = ukp; uim = uimkp;
    vkm = v0; v0 = vkp; vjm = vjmkp;
    wkm = w0; w0 = wkp; wipkm = wip; wjpkm = wjp;
    p0 = pkp;
    dzf_k = dzf_kp; rdzf_k = rdzf_kp; rdzc_km = rdzc_k;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace

// Tensor map of a ghosted field for boxes of bx x by x 1 doubles (tma.cuh: base one element before the (0,0,0) corner,
// or at element (1,0,0) when interior; both 16-byte aligned: the corner sits 15 doubles into a 256-byte aligned
// allocation, Field3).
bool tma_make_map(void *out128, const double *corner, const Grid3 &g, int bx, int by, bool interior) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)g.sy, (cuuint64_t)(g.n[1] + 2), (cuuint64_t)(g.n[2] + 2)};
  const cuuint64_t gstr[2] = {(cuuint64_t)g.sy * 8, (cuuint64_t)g.sz * 8};
  const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, estr[3] = {1, 1, 1};
  void *base = const_cast<double *>(corner) + (interior ? 1 : -1);
  if (reinterpret_cast<size_t>(base) & 15) return false;
  return fn(static_cast<CUtensorMap *>(out128), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the BX x BY tiles with halo of rk_mom_tma_kernel (interior = false) or the TX x TY tiles of an RK-history array
bool mom_tma_make_map(void *out128, const double *corner, const Grid3 &g, bool interior) {
  return tma_make_map(out128, corner, g, interior ? TX : BX, interior ? TY : BY, interior);
}

// returns false when this box / configuration is left to rk_mom_kernel (buoyancy term, no tensor maps)
bool launch_rk_mom_tma(Ctx &ctx, Box &b, int irk, double dt) {
  static const bool off = getenv("SNAC_MOM_TMA") && atoi(getenv("SNAC_MOM_TMA")) == 0;   // A/B switch against mom.cu
  if (off || ctx.prm.boussinesq || !b.tma_ok) return false;
  static const double RK[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};
  MomTmaArgs a;
  const int cur = ctx.par, nxt = 1 - ctx.par;
  for (int d = 0; d < 3; ++d) {
    a.n[d] = b.g.n[d];
    a.df[d] = b.df[d].p(); a.rdc[d] = b.rdc[d].p(); a.rdf[d] = b.rdf[d].p();
    a.bforce[d] = ctx.prm.bforce[d];
  }
  a.sy = b.g.sy; a.sz = b.g.sz;
  a.un = b.u[nxt].p; a.vn = b.v[nxt].p; a.wn = b.w[nxt].p;
  a.du = b.du.p; a.dv = b.dv.p; a.dw = b.dw.p;
  a.f1 = RK[irk - 1][0] * dt; a.f2 = RK[irk - 1][1] * dt; a.f12 = a.f1 + a.f2;
  a.visc = ctx.prm.visc;
  const int tiles = ((b.g.n[0] + TX - 1) / TX) * ((b.g.n[1] + TY - 1) / TY);
  // chunks of SNAC_MOM_KZ planes; shorter ones (down to 8 planes) when a small box would not fill the device otherwise
  int nchunk = std::max(1, std::max(b.g.n[2] / SNAC_MOM_KZ, std::min(b.g.n[2] / 8, (296 + tiles - 1) / tiles)));
  a.kz = (b.g.n[2] + nchunk - 1) / nchunk;
  nchunk = (b.g.n[2] + a.kz - 1) / a.kz;
  dim3 blk(TX, TY, 1), grd((b.g.n[0] + TX - 1) / TX, (b.g.n[1] + TY - 1) / TY, nchunk);
  const bool imp = ctx.prm.impdiff, first = (a.f2 == 0.0);
  const size_t smem = (size_t)NS * (first ? slot_doubles<true>() : slot_doubles<false>()) * sizeof(double);
  const CUtensorMap *th[3] = {reinterpret_cast<const CUtensorMap *>(b.tmap_hist[0]), reinterpret_cast<const CUtensorMap *>(b.tmap_hist[1]),
                              reinterpret_cast<const CUtensorMap *>(b.tmap_hist[2])};
  const CUtensorMap *tu = reinterpret_cast<const CUtensorMap *>(b.tmap[cur][0]), *tv = reinterpret_cast<const CUtensorMap *>(b.tmap[cur][1]);
  const CUtensorMap *tw = reinterpret_cast<const CUtensorMap *>(b.tmap[cur][2]), *tp = reinterpret_cast<const CUtensorMap *>(b.tmap[0][3]);
#define LAUNCH(I, F)                                                                                                           \
  {                                                                                                                            \
    static bool attr = false;                                                                                                  \
    if (!attr) {                                                                                                               \
      SNAC_CUDA(cudaFuncSetAttribute(rk_mom_tma_kernel<I, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
      attr = true;                                                                                                             \
    }                                                                                                                          \
    rk_mom_tma_kernel<I, F><<<grd, blk, smem, ctx.stream>>>(*tu, *tv, *tw, *tp, *th[0], *th[1], *th[2], a);                                            \
  }
  if (imp) { if (first) LAUNCH(true, true) else LAUNCH(true, false) }
  else     { if (first) LAUNCH(false, true) else LAUNCH(false, false) }
#undef LAUNCH
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
  return true;
}

}  // namespace snac
