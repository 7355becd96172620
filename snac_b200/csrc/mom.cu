// K1: fused momentum right-hand side + low-storage RK3 update.
//
// Replaces rk_mom (reference src/rk.f90:7-128) and the twelve sweeps it calls:
// mom{x,y,z}_a_vv (src/mom.f90:112-219), mom{x,y,z}_d (:221-327), mom{x,y,z}_p (:328-386).
// One thread per interior cell reads u,v,w (9 points each), p (4 points) and the three RK history
// arrays, and writes the three new velocities (into the other velocity buffer: the stencil reads
// neighbours, so the update cannot be done in place) and the three histories.
// Algorithmic traffic: 7 reads + 6 writes of FP64 = 104 B per cell (80 B when irk = 1, f2 = 0).
#include "common.h"

namespace snac {

struct MomArgs {
  BoxGeom g;
  const double *u, *v, *w, *p;
  double *un, *vn, *wn;
  double *du, *dv, *dw;
  const double *s;   // first scalar (Boussinesq) or nullptr
  double f1, f2, f12, visc;
  double bforce[3];
  double gb[3];      // gacc*beta
};

// Block shape and register budget, measured on TGV-256 (B200): 64x4 threads / 80 registers 1.40 ms per
// step, 32x8 and 128x2 the same, 64x2 / 80 registers 1.33 ms, 64x2 / 72 registers (7 CTAs per SM,
// 36 bytes of spills) 1.31 ms, 64x1 / 72 registers 1.29 ms; 64x4 squeezed into 64 registers 1.58 ms.
#ifndef SNAC_MOM_MINB
#define SNAC_MOM_MINB 14
#endif
#ifndef SNAC_MOM_BX
#define SNAC_MOM_BX 64
#define SNAC_MOM_BY 1
#endif
template <bool IMPDIFF, bool BUOY, bool FIRST>
__global__ void __launch_bounds__(SNAC_MOM_BX * SNAC_MOM_BY, SNAC_MOM_MINB) rk_mom_kernel(const MomArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
  const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
  const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
  if (i > a.g.n[0] || j > a.g.n[1] || k > a.g.n[2]) return;
  const long sy = a.g.sy, sz = a.g.sz;
  const long c = i + sy * j + sz * k;

  const double dxf_i = __ldg(a.g.df[0] + i), dxf_ip = __ldg(a.g.df[0] + i + 1);
  const double dyf_j = __ldg(a.g.df[1] + j), dyf_jp = __ldg(a.g.df[1] + j + 1);
  const double dzf_k = __ldg(a.g.df[2] + k), dzf_kp = __ldg(a.g.df[2] + k + 1);
  const double rdxf_i = __ldg(a.g.rdf[0] + i), rdxf_ip = __ldg(a.g.rdf[0] + i + 1);
  const double rdyf_j = __ldg(a.g.rdf[1] + j), rdyf_jp = __ldg(a.g.rdf[1] + j + 1);
  const double rdzf_k = __ldg(a.g.rdf[2] + k), rdzf_kp = __ldg(a.g.rdf[2] + k + 1);
  const double rdxc_i = __ldg(a.g.rdc[0] + i), rdxc_im = __ldg(a.g.rdc[0] + i - 1);
  const double rdyc_j = __ldg(a.g.rdc[1] + j), rdyc_jm = __ldg(a.g.rdc[1] + j - 1);
  const double rdzc_k = __ldg(a.g.rdc[2] + k), rdzc_km = __ldg(a.g.rdc[2] + k - 1);

  const double *u = a.u + c, *v = a.v + c, *w = a.w + c, *p = a.p + c;
  const double u0 = u[0], uip = u[1], uim = u[-1], ujp = u[sy], ujm = u[-sy], ukp = u[sz], ukm = u[-sz];
  const double uimjp = u[sy - 1], uimkp = u[sz - 1];
  const double v0 = v[0], vip = v[1], vim = v[-1], vjp = v[sy], vjm = v[-sy], vkp = v[sz], vkm = v[-sz];
  const double vipjm = v[1 - sy], vjmkp = v[sz - sy];
  const double w0 = w[0], wip = w[1], wim = w[-1], wjp = w[sy], wjm = w[-sy], wkp = w[sz], wkm = w[-sz];
  const double wipkm = w[1 - sz], wjpkm = w[sy - sz];
  const double p0 = p[0], pip = p[1], pjp = p[sy], pkp = p[sz];

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
  double px = -(pip - p0) * rdxc_i + a.bforce[0];
  double py = -(pjp - p0) * rdyc_j + a.bforce[1];
  double pz = -(pkp - p0) * rdzc_k + a.bforce[2];
  if (BUOY) {  // src/rk.f90:104-108
    const double *s = a.s + c;
    const double s0 = s[0];
    px -= a.gb[0] * 0.5 * (s[1] + s0);
    py -= a.gb[1] * 0.5 * (s[sy] + s0);
    pz -= a.gb[2] * 0.5 * (s[sz] + s0);
  }
  double rx, ry, rz;
  if (IMPDIFF) { rx = ax; ry = ay; rz = az; } else { rx = dx_ + ax; ry = dy_ + ay; rz = dz_ + az; }
  double un = u0 + a.f1 * rx, vn = v0 + a.f1 * ry, wn = w0 + a.f1 * rz;
  if (!FIRST) {
    un += a.f2 * a.du[c];
    vn += a.f2 * a.dv[c];
    wn += a.f2 * a.dw[c];
  }
  a.du[c] = rx; a.dv[c] = ry; a.dw[c] = rz;
  if (IMPDIFF) {  // + f12*d - 0.5*f12*d  (src/rk.f90:75-79,121-123)
    un += a.f12 * dx_; vn += a.f12 * dy_; wn += a.f12 * dz_;
  }
  un += a.f12 * px; vn += a.f12 * py; wn += a.f12 * pz;
  if (IMPDIFF) {
    un -= 0.5 * a.f12 * dx_; vn -= 0.5 * a.f12 * dy_; wn -= 0.5 * a.f12 * dz_;
  }
  a.un[c] = un; a.vn[c] = vn; a.wn[c] = wn;
}

static const double RK[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};

void launch_rk_mom(Ctx &ctx, Box &b, int irk, double dt) {
  MomArgs a;
  a.g = b.geom();
  const int cur = ctx.par, nxt = 1 - ctx.par;
  a.u = b.u[cur].p; a.v = b.v[cur].p; a.w = b.w[cur].p; a.p = b.p.p;
  a.un = b.u[nxt].p; a.vn = b.v[nxt].p; a.wn = b.w[nxt].p;
  a.du = b.du.p; a.dv = b.dv.p; a.dw = b.dw.p;
  a.s = ctx.prm.boussinesq ? b.s[0].p : nullptr;
  a.f1 = RK[irk - 1][0] * dt;
  a.f2 = RK[irk - 1][1] * dt;
  a.f12 = a.f1 + a.f2;
  a.visc = ctx.prm.visc;
  for (int d = 0; d < 3; ++d) {
    a.bforce[d] = ctx.prm.bforce[d];
    a.gb[d] = ctx.prm.gacc[d] * ctx.prm.beta;
  }
  dim3 blk(SNAC_MOM_BX, SNAC_MOM_BY, 1);
  if (b.g.n[0] <= 8) blk = dim3(8, 8, SNAC_MOM_BX * SNAC_MOM_BY / 64);
  else if (b.g.n[0] <= 32) blk = dim3(32, SNAC_MOM_BX * SNAC_MOM_BY / 32, 1);
  dim3 grd((b.g.n[0] + blk.x - 1) / blk.x, (b.g.n[1] + blk.y - 1) / blk.y, (b.g.n[2] + blk.z - 1) / blk.z);
  const bool imp = ctx.prm.impdiff, buo = ctx.prm.boussinesq, first = (a.f2 == 0.0);
#define LAUNCH(I, B, F) rk_mom_kernel<I, B, F><<<grd, blk, 0, ctx.stream>>>(a)
  if (imp) {
    if (buo) { if (first) LAUNCH(true, true, true); else LAUNCH(true, true, false); }
    else     { if (first) LAUNCH(true, false, true); else LAUNCH(true, false, false); }
  } else {
    if (buo) { if (first) LAUNCH(false, true, true); else LAUNCH(false, true, false); }
    else     { if (first) LAUNCH(false, false, true); else LAUNCH(false, false, false); }
  }
#undef LAUNCH
  ctx.launches++;
  SNAC_CUDA(cudaGetLastError());
}

}  // namespace snac
