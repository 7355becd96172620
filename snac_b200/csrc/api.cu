// C ABI of libsnacb200 (include/snac_b200.h) and the orchestration of one RK3 sub-step
// (reference src/main.f90:906-1033).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"

struct snac_handle {
  snac::Ctx ctx;
};

namespace snac {

static std::string g_create_error;

BoxGeom Box::geom() const {
  BoxGeom q;
  for (int d = 0; d < 3; ++d) {
    q.n[d] = g.n[d];
    q.dc[d] = dc[d].p(); q.df[d] = df[d].p(); q.rdc[d] = rdc[d].p(); q.rdf[d] = rdf[d].p();
  }
  q.sy = g.sy; q.sz = g.sz;
  return q;
}

Box &Ctx::box_of_rank(int rank) {
  if (rank < 0 || rank >= (int)local_of_rank.size() || local_of_rank[rank] < 0)
    throw Error("box " + std::to_string(rank) + " is not owned by this process");
  return *boxes[local_of_rank[rank]];
}

// set_bc (src/bound.f90:108-248) -> one face descriptor; returns false for 'F' (nothing to do)
static bool make_face_op(FaceOp &op, double *p, const Grid3 &g, int idir, int ib, char ctype, bool centered,
                         double rvalue, double dr) {
  if (ctype != 'D' && ctype != 'N') return false;
  double factor = rvalue, sgn = 0.0;
  if (ctype == 'D' && centered) { factor = 2.0 * factor; sgn = -1.0; }
  if (ctype == 'N') { factor = (ib == 0 ? -dr : dr) * factor; sgn = 1.0; }
  const int a = idir == 0 ? 1 : 0, b = idir == 2 ? 1 : 2;  // in-plane directions, lower one fastest
  op.p = p;
  op.sd = g.st(idir); op.sa = g.st(a); op.sb = g.st(b);
  op.n = g.n[idir]; op.na = g.n[a] + 2; op.nb = g.n[b] + 2;
  op.factor = factor; op.sgn = sgn;
  if (centered) op.kind = ib == 0 ? FACE_CEN_LO : FACE_CEN_HI;
  else if (ctype == 'D') op.kind = ib == 0 ? FACE_STAG_D_LO : FACE_STAG_D_HI;
  else op.kind = ib == 0 ? FACE_STAG_N_LO : FACE_STAG_N_HI;
  return true;
}

static void build_vel_bc(Ctx &ctx, int par, bool is_correc, FaceFill &ff) {
  std::vector<FaceOp> ops[3];
  for (auto &bp : ctx.boxes) {
    Box &b = *bp;
    double *fld[3] = {b.u[par].p, b.v[par].p, b.w[par].p};
    for (int idir = 0; idir < 3; ++idir)
      for (int ib = 0; ib < 2; ++ib) {
        if (!b.info.is_bound[idir][ib]) continue;
        const int q = ib == 0 ? 0 : b.g.n[idir];  // d?(lo-1) or d?(hi)
        for (int c = 0; c < 3; ++c) {
          FaceOp op;
          const bool normal = (c == idir);
          if (normal && is_correc) continue;  // src/bound.f90:37,42,48,53,59,64
          const double dr = normal ? b.h_df[idir][q] : b.h_dc[idir][q];
          if (make_face_op(op, fld[c], b.g, idir, ib, b.info.cbcvel[c][idir][ib], !normal, b.info.bcvel[c][idir][ib], dr))
            ops[idir].push_back(op);
        }
      }
  }
  ff.upload(ops);
}

static void build_cen_bc(Ctx &ctx, FaceFill &ff, double *(*get)(Box &, int), int arg, char (*cbc_of)(Box &, int, int, int),
                         double (*bc_of)(Box &, int, int, int)) {
  std::vector<FaceOp> ops[3];
  for (auto &bp : ctx.boxes) {
    Box &b = *bp;
    for (int idir = 0; idir < 3; ++idir)
      for (int ib = 0; ib < 2; ++ib) {
        if (!b.info.is_bound[idir][ib]) continue;
        const int q = ib == 0 ? 0 : b.g.n[idir];
        FaceOp op;
        if (make_face_op(op, get(b, arg), b.g, idir, ib, cbc_of(b, arg, idir, ib), true, bc_of(b, arg, idir, ib), b.h_dc[idir][q]))
          ops[idir].push_back(op);
      }
  }
  ff.upload(ops);
}

static void commit(Ctx &ctx) {
  if (ctx.committed) throw Error("snac_commit called twice");
  if (ctx.boxes.empty()) throw Error("snac_commit: no boxes were added");
  const int fa = ctx.prm.fft_axis - 1;
  for (auto &bp : ctx.boxes) {
    Box &b = *bp;
    for (int d = 0; d < 3; ++d) {
      const int n = b.g.n[d];
      std::vector<double> rc(n + 2), rf(n + 2), wq(n + 2), gq(n + 2);
      for (int i = 0; i <= n + 1; ++i) {
        rc[i] = 1.0 / b.h_dc[d][i];
        rf[i] = 1.0 / b.h_df[d][i];
        wq[i] = (d == fa) ? 1.0 : b.h_df[d][i];
        gq[i] = (d == fa) ? 1.0 / b.h_df[d][i] : 1.0;
      }
      b.dc[d].upload(b.h_dc[d]); b.df[d].upload(b.h_df[d]);
      b.rdc[d].upload(rc); b.rdf[d].upload(rf);
      b.wq[d].upload(wq); b.gq[d].upload(gq);
    }
    for (int q = 0; q < 2; ++q) { b.u[q].alloc(b.g); b.v[q].alloc(b.g); b.w[q].alloc(b.g); }
    b.p.alloc(b.g); b.rhs.alloc(b.g);
    b.pp_is_po = (ctx.prm.fft_axis == 0);
    if (b.pp_is_po) b.po.alloc(b.g);
    else { b.pp.alloc(b.g); b.stage.alloc(b.g); }
    b.du.alloc(b.g); b.dv.alloc(b.g); b.dw.alloc(b.g);
    for (int s = 0; s < ctx.prm.nscal; ++s) { b.s[s].alloc(b.g); b.ds[s].alloc(b.g); }
    compute_bc_rhs(ctx, b);
  }
  const int nl = (int)ctx.boxes.size();
  // halo exchanges
  for (int par = 0; par < 2; ++par) {
    std::vector<ExField> ef;
    for (int f = 0; f < 3; ++f)
      for (int lb = 0; lb < nl; ++lb) {
        Box &b = *ctx.boxes[lb];
        ef.push_back(ExField{lb, f == 0 ? b.u[par].p : (f == 1 ? b.v[par].p : b.w[par].p), b.g});
      }
    ctx.ex_uvw[par].build(ctx, ef);
    build_vel_bc(ctx, par, false, ctx.bc_uvw[par][0]);
    build_vel_bc(ctx, par, true, ctx.bc_uvw[par][1]);
  }
  {
    std::vector<ExField> ep, epp;
    for (int lb = 0; lb < nl; ++lb) {
      Box &b = *ctx.boxes[lb];
      ep.push_back(ExField{lb, b.p.p, b.g});
      epp.push_back(ExField{lb, b.pp_ptr(), b.g});
    }
    ctx.ex_p.build(ctx, ep);
    ctx.ex_pp.build(ctx, epp);
    auto cbc = [](Box &b, int, int d, int ib) -> char { return b.info.cbcpre[d][ib]; };
    auto bc = [](Box &b, int, int d, int ib) -> double { return b.info.bcpre[d][ib]; };
    build_cen_bc(ctx, ctx.bc_p, [](Box &b, int) { return b.p.p; }, 0, cbc, bc);
    build_cen_bc(ctx, ctx.bc_pp, [](Box &b, int) { return b.pp_ptr(); }, 0, cbc, bc);
  }
  for (int s = 0; s < ctx.prm.nscal; ++s) {
    std::vector<ExField> es;
    for (int lb = 0; lb < nl; ++lb) es.push_back(ExField{lb, ctx.boxes[lb]->s[s].p, ctx.boxes[lb]->g});
    ctx.ex_s[s].build(ctx, es);
    auto cbc = [](Box &b, int s_, int d, int ib) -> char { return b.info.cbcscal[s_][d][ib]; };
    auto bc = [](Box &b, int s_, int d, int ib) -> double { return b.info.bcscal[s_][d][ib]; };
    build_cen_bc(ctx, ctx.bc_s[s], [](Box &b, int s_) { return b.s[s_].p; }, s, cbc, bc);
  }
  // transforms (fftini for the pressure, src/main.f90:735) and the Poisson solver
  if (ctx.prm.fft_axis > 0) {
    Box &b0 = *ctx.boxes[0];
    const char bc[2] = {b0.info.cbcpre[fa][0], b0.info.cbcpre[fa][1]};
    for (auto &bp : ctx.boxes)
      if (bp->g.n[fa] != b0.g.n[fa] || bp->info.cbcpre[fa][0] != bc[0] || bp->info.cbcpre[fa][1] != bc[1])
        throw Error("all boxes must have the same size and boundary conditions along the transform axis (src/sanity.f90:66-100)");
    r2r_plan_init(ctx.plan_p, fa, b0.g.n[fa], bc, true);
  }
  {  // the pressure problem as the solver sees it: canonical (permuted) layout when batched
    const bool batched = ctx.prm.fft_axis > 0;
    MGSolver &S = ctx.psolver;
    if (batched) {
      int q = 0;
      for (int d = 0; d < 3; ++d) if (d != fa) S.perm[q++] = d;
      S.perm[2] = fa;
    }
    std::vector<SBox> sb(nl);
    for (int lb = 0; lb < nl; ++lb) {
      Box &b = *ctx.boxes[lb];
      SBox &s = sb[lb];
      for (int sd = 0; sd < 3; ++sd) {
        const int d = S.perm[sd];
        s.n[sd] = b.g.n[d];
        s.d1[sd] = b.h_dc[d];
        s.d2[sd] = b.h_df[d];
        for (int ib = 0; ib < 2; ++ib) {
          s.nb[sd][ib] = b.info.nb[d][ib];
          s.phys[sd][ib] = b.info.is_bound[d][ib] && ctx.prm.periods[d] == 0;
          s.sgn[sd][ib] = 0.0;
          if (s.phys[sd][ib] && !(batched && sd == 2)) {   // src/solver.f90:71-94 (centred unknowns)
            const char c = b.info.cbcpre[d][ib];
            if (c == 'N') s.sgn[sd][ib] = 1.0;
            else if (c == 'D') s.sgn[sd][ib] = -1.0;
            else throw Error("pressure boundary condition on a physical face must be 'D' or 'N'");
          }
        }
      }
      if (batched) {  // eigenvalues / dl^2 (src/fft.f90:151-189, src/solver.f90:1124-1138)
        const char bc[2] = {b.info.cbcpre[fa][0], b.info.cbcpre[fa][1]};
        std::vector<double> l = r2r_eigenvalues(b.g.n[fa], bc, true);
        const double dl = b.info.dl[fa][0];
        s.lam.assign(l.size() + 2, 0.0);
        for (size_t k = 0; k < l.size(); ++k) s.lam[k + 1] = l[k] / (dl * dl);
      } else {
        s.x = b.po.p;
        s.b = b.rhs.p;
      }
    }
    S.setup(ctx, std::move(sb), batched);
  }
  ctx.red.alloc(8);
  SNAC_CUDA(cudaMallocHost(&ctx.h_red, 8 * sizeof(double)));
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
  ctx.committed = true;
}

static Field3 *field_of(Ctx &ctx, Box &b, int field, bool &ghost_alias_pp) {
  ghost_alias_pp = false;
  switch (field) {
    case SNAC_FIELD_U: return &b.u[ctx.par];
    case SNAC_FIELD_V: return &b.v[ctx.par];
    case SNAC_FIELD_W: return &b.w[ctx.par];
    case SNAC_FIELD_P: return &b.p;
    case SNAC_FIELD_PP: return b.pp_is_po ? &b.po : &b.pp;
    case SNAC_FIELD_RHS: return &b.rhs;
    case SNAC_FIELD_PO: return &b.po;
    case SNAC_FIELD_DUDTRKO: return &b.du;
    case SNAC_FIELD_DVDTRKO: return &b.dv;
    case SNAC_FIELD_DWDTRKO: return &b.dw;
  }
  if (field >= SNAC_FIELD_S0 && field < SNAC_FIELD_S0 + ctx.prm.nscal) return &b.s[field - SNAC_FIELD_S0];
  if (field >= SNAC_FIELD_DSDTRKO0 && field < SNAC_FIELD_DSDTRKO0 + ctx.prm.nscal) return &b.ds[field - SNAC_FIELD_DSDTRKO0];
  throw Error("unknown field id " + std::to_string(field));
}

// strides of the physical directions in the solver's canonical array of local box lb
static void solver_strides(Ctx &ctx, int lb, long st[3]) {
  const MGSolver &S = ctx.psolver;
  const Grid3 &g = S.grid0(lb);
  for (int sd = 0; sd < 3; ++sd) st[S.perm[sd]] = g.st(sd);
}

__global__ void permute_copy_kernel(double *dst, long d0, long d1, long d2, const double *src, long s0, long s1, long s2,
                                    int n0, int n1, int n2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i < n0 && j < n1 && k < n2) dst[i * d0 + j * d1 + k * d2] = src[i * s0 + j * s1 + k * s2];
}

static void transfer(Ctx &ctx, int rank, int field, double *host, bool up) {
  Box &b = ctx.box_of_rank(rank);
  const int lb = ctx.local_of_rank[rank];
  const bool batched = ctx.prm.fft_axis > 0;
  // spectral arrays live in the solver's canonical layout: go through a staging array
  const bool spectral = batched && (field == SNAC_FIELD_PO || field == SNAC_FIELD_RHS_HAT);
  Field3 *f;
  bool dummy;
  if (spectral) f = &b.stage;
  else if (field == SNAC_FIELD_RHS_HAT) f = &b.rhs;
  else f = field_of(ctx, b, field, dummy);
  long sst[3] = {0, 0, 0};
  double *sp = nullptr;
  long pint = 1 + b.g.sy + b.g.sz, sint = 0;
  dim3 blk(32, 8, 1), grd((b.g.n[0] + 31) / 32, (b.g.n[1] + 7) / 8, b.g.n[2]);
  if (spectral) {
    solver_strides(ctx, lb, sst);
    const Grid3 &sg = ctx.psolver.grid0(lb);
    sint = 1 + sg.sy + sg.sz;
    sp = field == SNAC_FIELD_PO ? ctx.psolver.xcur(lb) : ctx.psolver.bcur(lb);
    if (!up) {
      SNAC_CUDA(cudaMemsetAsync(b.stage.buf.raw, 0, b.stage.buf.elems * sizeof(double), ctx.stream));
      permute_copy_kernel<<<grd, blk, 0, ctx.stream>>>(b.stage.p + pint, 1, b.g.sy, b.g.sz, sp + sint, sst[0], sst[1], sst[2],
                                                       b.g.n[0], b.g.n[1], b.g.n[2]);
    }
  }
  cudaMemcpy3DParms prm;
  std::memset(&prm, 0, sizeof(prm));
  const size_t w = (size_t)(b.g.n[0] + 2), h = (size_t)(b.g.n[1] + 2), dpt = (size_t)(b.g.n[2] + 2);
  cudaPitchedPtr hp = make_cudaPitchedPtr(host, w * sizeof(double), w, h);
  cudaPitchedPtr dp = make_cudaPitchedPtr(f->p, b.g.sy * sizeof(double), w, h);
  prm.srcPtr = up ? hp : dp;
  prm.dstPtr = up ? dp : hp;
  prm.extent = make_cudaExtent(w * sizeof(double), h, dpt);
  prm.kind = up ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
  SNAC_CUDA(cudaMemcpy3DAsync(&prm, ctx.stream));
  if (spectral && up)
    permute_copy_kernel<<<grd, blk, 0, ctx.stream>>>(sp + sint, sst[0], sst[1], sst[2], b.stage.p + pint, 1, b.g.sy, b.g.sz,
                                                     b.g.n[0], b.g.n[1], b.g.n[2]);
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
}

// ---- pieces of the sub-step ------------------------------------------------------------------
static void do_rk_mom(Ctx &ctx, int irk, double dt) {
  if (irk < 1 || irk > 3) throw Error("rk_mom: irk must be 1, 2 or 3");
  ProfScope ps(ctx, PC_RK_MOM);
  for (auto &b : ctx.boxes) launch_rk_mom(ctx, *b, irk, dt);
  ctx.par = 1 - ctx.par;
}
static void do_bounduvw(Ctx &ctx, bool is_correc) {
  ProfScope ps(ctx, PC_HALO_BC);
  ctx.ex_uvw[ctx.par].run_all(ctx);
  ctx.bc_uvw[ctx.par][is_correc ? 1 : 0].run(ctx);
}
static void do_boundp(Ctx &ctx, int field) {
  ProfScope ps(ctx, PC_HALO_BC);
  if (field == SNAC_FIELD_P) { ctx.ex_p.run_all(ctx); ctx.bc_p.run(ctx); }
  else if (field == SNAC_FIELD_PP) { ctx.ex_pp.run_all(ctx); ctx.bc_pp.run(ctx); }
  else if (field >= SNAC_FIELD_S0 && field < SNAC_FIELD_S0 + ctx.prm.nscal) {
    ctx.ex_s[field - SNAC_FIELD_S0].run_all(ctx);
    ctx.bc_s[field - SNAC_FIELD_S0].run(ctx);
  } else throw Error("boundp: field must be P, PP or a scalar");
}
static void do_fft(Ctx &ctx, int dir) {
  if (ctx.prm.fft_axis == 0) throw Error("fft: the handle was created with fft_axis = 0");
  ProfScope ps(ctx, dir > 0 ? PC_FFT_FWD : PC_FFT_BWD);
  const MGSolver &S = ctx.psolver;
  for (int lb = 0; lb < (int)ctx.boxes.size(); ++lb) {
    Box &b = *ctx.boxes[lb];
    long pst[3] = {1, b.g.sy, b.g.sz}, sst[3];
    solver_strides(ctx, lb, sst);
    const Grid3 &sg = S.grid0(lb);
    const long pint = 1 + b.g.sy + b.g.sz, sint = 1 + sg.sy + sg.sz;   // first interior element
    if (dir > 0) r2r_execute(ctx, ctx.plan_p, +1, b.g.n, b.rhs.p + pint, pst, S.bcur(lb) + sint, sst, 1.0);
    else r2r_execute(ctx, ctx.plan_p, -1, b.g.n, S.xcur(lb) + sint, sst, b.pp.p + pint, pst, ctx.plan_p.normfft);
  }
}
static void do_solve(Ctx &ctx) { ctx.psolver.solve(ctx); }

static void do_chk(Ctx &ctx, bool want_dt, double *out3) {
  SNAC_CUDA(cudaMemsetAsync(ctx.red.raw, 0, 8 * sizeof(double), ctx.stream));
  {
    ProfScope ps(ctx, PC_CHK);
    for (auto &b : ctx.boxes) launch_chk(ctx, *b, ctx.red.raw);
  }
  SNAC_CUDA(cudaMemcpyAsync(ctx.h_red, ctx.red.raw, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
  (void)want_dt;
  for (int q = 0; q < 3; ++q) out3[q] = ctx.h_red[q];
}

static void substep(Ctx &ctx, int irk, double dt) {
  static const double RK[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};
  const double dtrk = (RK[irk - 1][0] + RK[irk - 1][1]) * dt;
  const double alpha = -ctx.prm.visc * dtrk / 2.0;
  if (ctx.prm.nscal > 0 || ctx.prm.impdiff)
    throw Error("snac_step: scalar transport / implicit diffusion are not available in this build");
  do_rk_mom(ctx, irk, dt);
  do_bounduvw(ctx, false);
  {
    ProfScope ps(ctx, PC_FILLPS);
    for (auto &b : ctx.boxes) launch_fillps_bc(ctx, *b, dtrk);
  }
  if (ctx.prm.fft_axis > 0) {
    do_fft(ctx, +1);
    do_solve(ctx);
    do_fft(ctx, -1);
  } else {
    do_solve(ctx);
  }
  do_boundp(ctx, SNAC_FIELD_PP);
  {
    ProfScope ps(ctx, PC_CORREC);
    for (auto &b : ctx.boxes) launch_correc_updt(ctx, *b, dtrk, alpha, true, true);
  }
  do_bounduvw(ctx, true);
  do_boundp(ctx, SNAC_FIELD_P);
}

}  // namespace snac

using namespace snac;

#define SNAC_TRY(h, ...)                          \
  if (!(h)) return -2;                            \
  try {                                           \
    __VA_ARGS__;                                  \
    return 0;                                     \
  } catch (const std::exception &e) {             \
    (h)->ctx.err = e.what();                      \
    return -1;                                    \
  }

extern "C" {

int snac_create(const snac_params *params, snac_handle **out) {
  if (!params || !out) return -2;
  *out = nullptr;
  try {
    if (params->abi_version != SNAC_ABI_VERSION) throw Error("snac_create: ABI version mismatch");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(std::string("snac_create: no CUDA device available (") + cudaGetErrorString(e) + "); there is no CPU path");
    if (params->nscal < 0 || params->nscal > SNAC_MAX_SCAL) throw Error("snac_create: nscal out of range");
    if (params->fft_axis < 0 || params->fft_axis > 3) throw Error("snac_create: fft_axis must be 0..3");
    SNAC_CUDA(cudaSetDevice(params->device));
    std::unique_ptr<snac_handle> h(new snac_handle);
    Ctx &ctx = h->ctx;
    ctx.prm = *params;
    ctx.owner.assign(params->nrank, 0);
    if (params->rank_owner) ctx.owner.assign(params->rank_owner, params->rank_owner + params->nrank);
    ctx.prm.rank_owner = nullptr;
    ctx.local_of_rank.assign(params->nrank, -1);
    SNAC_CUDA(cudaStreamCreateWithFlags(&ctx.stream, cudaStreamNonBlocking));
    SNAC_CUDA(cudaEventCreate(&ctx.prof_ev[0]));
    SNAC_CUDA(cudaEventCreate(&ctx.prof_ev[1]));
    SNAC_CUDA(cudaEventCreate(&ctx.timer_ev[0]));
    SNAC_CUDA(cudaEventCreate(&ctx.timer_ev[1]));
    *out = h.release();
    return 0;
  } catch (const std::exception &e) {
    g_create_error = e.what();
    return -1;
  }
}

const char *snac_last_error(const snac_handle *h) { return h ? h->ctx.err.c_str() : g_create_error.c_str(); }

int snac_add_box(snac_handle *h, const snac_box *box) {
  SNAC_TRY(h, {
    Ctx &ctx = h->ctx;
    if (ctx.committed) throw Error("snac_add_box after snac_commit");
    if (!box || box->rank < 0 || box->rank >= ctx.prm.nrank) throw Error("snac_add_box: bad box id");
    if (ctx.owner[box->rank] != ctx.prm.myproc) throw Error("snac_add_box: box is owned by another process");
    if (ctx.local_of_rank[box->rank] >= 0) throw Error("snac_add_box: box added twice");
    std::unique_ptr<Box> b(new Box);
    b->info = *box;
    int n[3];
    for (int d = 0; d < 3; ++d) {
      n[d] = box->hi[d] - box->lo[d] + 1;
      if (n[d] < 1) throw Error("snac_add_box: empty box");
    }
    b->g = Grid3::make(n[0], n[1], n[2]);
    const double *dc[3] = {box->dxc, box->dyc, box->dzc}, *df[3] = {box->dxf, box->dyf, box->dzf};
    for (int d = 0; d < 3; ++d) {
      if (!dc[d] || !df[d]) throw Error("snac_add_box: missing grid arrays");
      b->h_dc[d].assign(dc[d], dc[d] + n[d] + 2);
      b->h_df[d].assign(df[d], df[d] + n[d] + 2);
    }
    b->info.dxc = b->info.dxf = b->info.dyc = b->info.dyf = b->info.dzc = b->info.dzf = nullptr;
    ctx.local_of_rank[box->rank] = (int)ctx.boxes.size();
    ctx.boxes.push_back(std::move(b));
  })
}

int snac_commit(snac_handle *h) { SNAC_TRY(h, commit(h->ctx)) }

int snac_finalize(snac_handle *h) {
  if (!h) return -2;
  cudaStreamSynchronize(h->ctx.stream);
  if (h->ctx.h_red) cudaFreeHost(h->ctx.h_red);
  cudaStream_t s = h->ctx.stream;
  delete h;
  if (s) cudaStreamDestroy(s);
  return 0;
}

int snac_upload(snac_handle *h, int rank, int field, const double *host) {
  SNAC_TRY(h, transfer(h->ctx, rank, field, const_cast<double *>(host), true))
}
int snac_download(snac_handle *h, int rank, int field, double *host) {
  SNAC_TRY(h, transfer(h->ctx, rank, field, host, false))
}

int snac_rk_mom(snac_handle *h, int irk, double dt) { SNAC_TRY(h, do_rk_mom(h->ctx, irk, dt)) }
int snac_rk_scal(snac_handle *h, int, double, int) {
  SNAC_TRY(h, throw Error("snac_rk_scal: scalar transport is not available in this build"))
}
int snac_bounduvw(snac_handle *h, int is_correc) { SNAC_TRY(h, do_bounduvw(h->ctx, is_correc != 0)) }
int snac_boundp(snac_handle *h, int field) { SNAC_TRY(h, do_boundp(h->ctx, field)) }
int snac_fillps(snac_handle *h, double dtrk) {
  SNAC_TRY(h, for (auto &b : h->ctx.boxes) launch_fillps(h->ctx, *b, dtrk))
}
int snac_updt_rhs(snac_handle *h) { SNAC_TRY(h, for (auto &b : h->ctx.boxes) launch_updt_rhs(h->ctx, *b)) }
int snac_fft(snac_handle *h, int dir) { SNAC_TRY(h, do_fft(h->ctx, dir)) }
int snac_solve_n_helmholtz_2d(snac_handle *h) {
  SNAC_TRY(h, {
    if (h->ctx.prm.fft_axis == 0) throw Error("solve_n_helmholtz_2d needs fft_axis > 0");
    do_solve(h->ctx);
  })
}
int snac_solve_helmholtz(snac_handle *h) {
  SNAC_TRY(h, {
    if (h->ctx.prm.fft_axis != 0) throw Error("solve_helmholtz needs fft_axis = 0");
    do_solve(h->ctx);
  })
}
int snac_correc(snac_handle *h, double dtrk) {
  SNAC_TRY(h, for (auto &b : h->ctx.boxes) launch_correc_updt(h->ctx, *b, dtrk, 0.0, true, false))
}
int snac_updt_pressure(snac_handle *h, double alpha) {
  SNAC_TRY(h, for (auto &b : h->ctx.boxes) launch_correc_updt(h->ctx, *b, 0.0, alpha, false, true))
}

int snac_chkdt(snac_handle *h, double *dt_cfl) {
  SNAC_TRY(h, {
    Ctx &ctx = h->ctx;
    double r[3];
    do_chk(ctx, true, r);
    double dti = r[0];
    if (dti == 0.0) dti = 1.0;
    double dtmax = std::sqrt(3.0) / dti;
    if (!ctx.prm.impdiff) {  // src/chkdt.f90:51-55
      double dlmin = 1e300;
      for (auto &b : ctx.boxes)
        for (int d = 0; d < 3; ++d)
          for (double v : b->h_df[d]) dlmin = std::min(dlmin, v);
      double alpha_max = 0.0;
      for (int s = 0; s < ctx.prm.nscal; ++s) alpha_max = std::max(alpha_max, 1.0 / ctx.prm.alphai[s]);
      const double visc = std::max(ctx.prm.visc, alpha_max);
      dtmax = std::min(dtmax, 1.65 / 12.0 / visc * dlmin * dlmin);
    }
    *dt_cfl = dtmax;
  })
}
int snac_chkdiv(snac_handle *h, double *divtot, double *divmax) {
  SNAC_TRY(h, {
    double r[3];
    do_chk(h->ctx, false, r);
    *divtot = r[1];
    *divmax = r[2];
  })
}

int snac_step(snac_handle *h, double dt) {
  SNAC_TRY(h, for (int irk = 1; irk <= 3; ++irk) substep(h->ctx, irk, dt))
}

int snac_solver_stats(snac_handle *h, int *iters, double *relres) {
  SNAC_TRY(h, {
    h->ctx.psolver.stats(h->ctx, iters, relres, nullptr);
  })
}
int snac_profile_enable(snac_handle *h, int on) {
  SNAC_TRY(h, {
    SNAC_CUDA(cudaStreamSynchronize(h->ctx.stream));
    h->ctx.prof_on = on != 0;
    if (on) for (int c = 0; c < PC_COUNT; ++c) { h->ctx.prof_ms[c] = 0.0; h->ctx.prof_n[c] = 0; }
  })
}
int snac_profile_get(snac_handle *h, int cls, double *ms, long long *launches) {
  SNAC_TRY(h, {
    if (cls < 0 || cls >= PC_COUNT) throw Error("snac_profile_get: class out of range");
    if (ms) *ms = h->ctx.prof_ms[cls];
    if (launches) *launches = h->ctx.prof_n[cls];
  })
}
const char *snac_profile_name(int cls) {
  static const char *names[PC_COUNT] = {"rk_mom", "halo_bc", "fillps", "fft_fwd", "mg_smooth_l0", "mg_down_l0",
                                        "mg_up_l0", "mg_coarse_levels", "fft_bwd", "correc_updt_pressure", "chkdt_chkdiv"};
  return (cls >= 0 && cls < PC_COUNT) ? names[cls] : nullptr;
}
int snac_timer_start(snac_handle *h) { SNAC_TRY(h, SNAC_CUDA(cudaEventRecord(h->ctx.timer_ev[0], h->ctx.stream))) }
int snac_timer_stop(snac_handle *h, double *ms) {
  SNAC_TRY(h, {
    SNAC_CUDA(cudaEventRecord(h->ctx.timer_ev[1], h->ctx.stream));
    SNAC_CUDA(cudaEventSynchronize(h->ctx.timer_ev[1]));
    float f = 0.f;
    SNAC_CUDA(cudaEventElapsedTime(&f, h->ctx.timer_ev[0], h->ctx.timer_ev[1]));
    if (ms) *ms = f;
  })
}
int snac_solver_mean_iters(snac_handle *h, double *mean_iters) {
  SNAC_TRY(h, h->ctx.psolver.stats(h->ctx, nullptr, nullptr, mean_iters))
}
long long snac_launch_count(const snac_handle *h) { return h ? h->ctx.launches : -1; }
int snac_sync(snac_handle *h) { SNAC_TRY(h, SNAC_CUDA(cudaStreamSynchronize(h->ctx.stream))) }
void *snac_stream(snac_handle *h) { return h ? (void *)h->ctx.stream : nullptr; }

}  // extern "C"
