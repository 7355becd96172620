// C ABI of libsnacb200 (include/snac_b200.h) and the orchestration of one RK3 sub-step
// (reference src/main.f90:906-1033).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"
#include "fft_radix.cuh"

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

// block-level description used by the batched solver: the boxes of a block are merged so that
// every 2-D problem lives on one array per block (the transform kernels scatter/gather the pieces)
struct BlockInfo {
  std::vector<int> members;   // ids of the blocks merged into this one (merge_blocks), first = id
  int id;
  int lo_g[3], hi_g[3];
  int nb[3][2];         // neighbour block or -1
  int face_rank[3][2];  // a box touching that face
  std::vector<double> dc[3], df[3];
};

static std::vector<BlockInfo> assemble_blocks(Ctx &ctx) {
  std::vector<BlockInfo> blocks;
  std::vector<int> index_of;   // block id -> index
  for (int r = 0; r < ctx.prm.nrank; ++r) {
    if (!ctx.world_set[r]) throw Error("the batched solver needs the description of every box of the run: add box " + std::to_string(r));
    const snac_box &w = ctx.world[r];
    if ((int)index_of.size() <= w.block) index_of.resize(w.block + 1, -1);
    if (index_of[w.block] < 0) {
      index_of[w.block] = (int)blocks.size();
      BlockInfo B;
      B.id = w.block;
      B.members.assign(1, w.block);
      for (int d = 0; d < 3; ++d) {
        B.lo_g[d] = w.lo[d]; B.hi_g[d] = w.hi[d];
        for (int ib = 0; ib < 2; ++ib) { B.nb[d][ib] = -1; B.face_rank[d][ib] = -1; }
      }
      blocks.push_back(B);
    }
    BlockInfo &B = blocks[index_of[w.block]];
    for (int d = 0; d < 3; ++d) { B.lo_g[d] = std::min(B.lo_g[d], w.lo[d]); B.hi_g[d] = std::max(B.hi_g[d], w.hi[d]); }
  }
  for (BlockInfo &B : blocks)
    for (int d = 0; d < 3; ++d) {
      const int N = B.hi_g[d] - B.lo_g[d] + 1;
      B.dc[d].assign(N + 2, 0.0);
      B.df[d].assign(N + 2, 0.0);
    }
  for (int r = 0; r < ctx.prm.nrank; ++r) {
    const snac_box &w = ctx.world[r];
    BlockInfo &B = blocks[index_of[w.block]];
    for (int d = 0; d < 3; ++d) {
      const int n = w.hi[d] - w.lo[d] + 1, o = w.lo[d] - B.lo_g[d];
      for (int q = 0; q <= n + 1; ++q) { B.dc[d][o + q] = ctx.world_dc[d][r][q]; B.df[d][o + q] = ctx.world_df[d][r][q]; }
      if (w.lo[d] == B.lo_g[d]) B.face_rank[d][0] = r;
      if (w.hi[d] == B.hi_g[d]) B.face_rank[d][1] = r;
    }
  }
  for (BlockInfo &B : blocks)
    for (int d = 0; d < 3; ++d)
      for (int ib = 0; ib < 2; ++ib) {
        const int fr = B.face_rank[d][ib];
        const int nbr = ctx.world[fr].nb[d][ib];
        B.nb[d][ib] = nbr < 0 ? -1 : index_of[ctx.world[nbr].block];
      }
  return blocks;
}

// Friend blocks whose union is one rectangle are one block as far as the 2-D solves are concerned: the matrix rows
// of init_matrix_2d (src/solver.f90:687-842) couple the cells across a friend face exactly as they couple cells
// inside a block (the halo of the neighbour carries the unknown, spacings come from bound_grid), so the per-wavenumber
// problem of e.g. the reference's 2 x 2-block duct (examples/_manuscript_turbulent_duct, periodic_duct) is the
// problem of one block cut differently.  Merged, its planes are whole in the slab layout: the plane-local
// multigrid kernels and the direct solver apply, on any number of GPUs, and no block-interface exchange is left in
// the solver.  Returns the blocks unchanged unless ALL of them tile their bounding box, the 1-D grids agree where
// blocks overlap in index space and every outer face carries one boundary condition.
static std::vector<BlockInfo> merge_blocks(Ctx &ctx, const std::vector<BlockInfo> &blocks) {
  const bool off = getenv("SNAC_MERGE_BLOCKS") && atoi(getenv("SNAC_MERGE_BLOCKS")) == 0;   // test hook: keep the general path
  if (blocks.size() < 2 || off) return blocks;
  BlockInfo M;
  M.id = blocks[0].id;
  for (const BlockInfo &B : blocks) { M.members.push_back(B.id); M.id = std::min(M.id, B.id); }
  double vol = 0.0;
  for (int d = 0; d < 3; ++d) { M.lo_g[d] = blocks[0].lo_g[d]; M.hi_g[d] = blocks[0].hi_g[d]; }
  for (const BlockInfo &B : blocks) {
    double v = 1.0;
    for (int d = 0; d < 3; ++d) {
      M.lo_g[d] = std::min(M.lo_g[d], B.lo_g[d]); M.hi_g[d] = std::max(M.hi_g[d], B.hi_g[d]);
      v *= B.hi_g[d] - B.lo_g[d] + 1;
    }
    vol += v;
  }
  double box = 1.0;
  for (int d = 0; d < 3; ++d) box *= M.hi_g[d] - M.lo_g[d] + 1;
  if (vol != box) return blocks;   // holes or overlaps: not a rectangle
  for (int d = 0; d < 3; ++d) {
    const int N = M.hi_g[d] - M.lo_g[d] + 1;
    M.dc[d].assign(N + 2, 0.0); M.df[d].assign(N + 2, 0.0);
    std::vector<char> set(N + 2, 0);
    for (const BlockInfo &B : blocks) {
      const int n = B.hi_g[d] - B.lo_g[d] + 1, o = B.lo_g[d] - M.lo_g[d];
      for (int q = 0; q <= n + 1; ++q) {
        // a block's ghost entries (q = 0, n+1) are only authoritative on the outer faces of the union
        const bool ghost = q == 0 || q == n + 1;
        if (ghost && !((q == 0 && o == 0) || (q == n + 1 && o + n == N))) {
          if (q == 0) { if (!set[o]) { M.dc[d][o] = B.dc[d][0]; } else if (std::fabs(M.dc[d][o] - B.dc[d][0]) > 1e-13 * std::fabs(B.dc[d][0])) return blocks; }
          continue;
        }
        if (set[o + q] && (std::fabs(M.dc[d][o + q] - B.dc[d][q]) > 1e-13 * std::fabs(B.dc[d][q]) ||
                           std::fabs(M.df[d][o + q] - B.df[d][q]) > 1e-13 * std::fabs(B.df[d][q])))
          return blocks;
        M.dc[d][o + q] = B.dc[d][q]; M.df[d][o + q] = B.df[d][q];
        set[o + q] = 1;
      }
    }
    for (int ib = 0; ib < 2; ++ib) {
      // all blocks on this outer face: same kind (physical with one condition, or periodic onto the union)
      M.face_rank[d][ib] = -1;
      int kind = -2;   // -1 physical, >= 0 periodic
      for (const BlockInfo &B : blocks) {
        const bool outer = ib == 0 ? B.lo_g[d] == M.lo_g[d] : B.hi_g[d] == M.hi_g[d];
        if (!outer) { if (B.nb[d][ib] < 0) return blocks; continue; }   // an inner face must have a friend
        const int fr = B.face_rank[d][ib];
        const int k = B.nb[d][ib] < 0 ? -1 : 0;
        if (kind == -2) { kind = k; M.face_rank[d][ib] = fr; }
        else if (kind != k) return blocks;
        else if (ctx.world[fr].dl[d][ib] != ctx.world[M.face_rank[d][ib]].dl[d][ib] ||
                 (k < 0 && (ctx.world[fr].cbcpre[d][ib] != ctx.world[M.face_rank[d][ib]].cbcpre[d][ib] ||
                            ctx.world[fr].bcpre[d][ib] != ctx.world[M.face_rank[d][ib]].bcpre[d][ib])))
          return blocks;   // (on 'F' faces bcpre holds the friend block's number, which differs from block to block)
      }
      M.nb[d][ib] = kind >= 0 ? 0 : -1;   // periodic wrap onto the merged block itself (index 0)
    }
  }
  return std::vector<BlockInfo>(1, M);
}

// Which wavenumbers does process `proc` solve?  Cyclic (q mod nproc == proc) when the number of
// processes is a power of two: the cycles a plane needs fall with its wavenumber, so consecutive
// chunks would leave the process holding the lowest ones working while the others wait.
// Returns the first wavenumber q0 and the count (as q1 - q0); wavenumber k of the slab is
// q0 + k * q_stride(ctx).
static bool q_is_cyclic(const Ctx &ctx) {
  const int np = ctx.comm.nproc;
  return np > 1 && (np & (np - 1)) == 0;
}
static int q_stride(const Ctx &ctx) { return q_is_cyclic(ctx) ? ctx.comm.nproc : 1; }
static void q_range(const Ctx &ctx, int nq, int proc, int &q0, int &q1) {
  if (q_is_cyclic(ctx)) {
    q0 = std::min(nq, proc);
    q1 = q0 + (nq > proc ? (nq - proc + ctx.comm.nproc - 1) / ctx.comm.nproc : 0);
    return;
  }
  const int chunk = (nq + ctx.comm.nproc - 1) / ctx.comm.nproc;
  q0 = std::min(nq, proc * chunk);
  q1 = std::min(nq, q0 + chunk);
}

static void alloc_field(Ctx &ctx, Field3 &f, const Grid3 &g, int id, int key) {
  if (ctx.comm.nproc > 1) f.adopt(ctx.comm.alloc(g.alloc_elems(), id, key));
  else f.alloc(g);
}

// Implicit diffusion (IMPDIFF): one Helmholtz problem per velocity component on its own staggered
// grid and one per scalar (src/main.f90:770-886).  The operators are strongly diagonally dominant
// (diagonal shift alphai = -2/(nu dtrk)), and the staggered unknowns do not coarsen by cell pairs,
// so they are solved on a single level with red-black Gauss-Seidel sweeps; the shifted diagonal is a
// kernel argument, so nothing is re-assembled when dt changes (the reference re-creates and
// re-sets-up its HYPRE solver on every call, src/solve_helmholtz.f90:85-86,134-135).
// With FFT_AXIS on, the reference reduces each of these problems along the transform axis as it does for the pressure
// (solve_impdiff_field_fft, src/solve_helmholtz.f90:32-100: scale, updt_rhs, transform, one 2-D solve per wavenumber,
// back); that is a second algorithm for the SAME linear system (the transform diagonalises the discrete operator of
// that direction, src/fft.f90:96-189), so the 3-D Gauss-Seidel solve below serves both builds: for a diagonal shift of
// 2/(nu dtrk) it contracts by orders of magnitude per sweep and moves 24 B per point and sweep, less than a
// transform pair + plane solves would (80 B per point).  The pressure keeps its transform path.
static double helm_bc_coef(const Box &b, int d, int ib, char c, double bc, bool centered, int nhi, const std::vector<double> &d1,
                           const std::vector<double> &d2, const int periods[3]) {
  if (!b.info.is_bound[d][ib] || periods[d] != 0) return 0.0;
  if (ib == 1 && nhi == 1) return 0.0;   // the reference loop visits i==lo only (ib=0)
  double factor = 0.0;
  const double dl = centered ? b.info.dl[d][ib] : b.info.dlf[d][ib];
  if (c == 'N') factor = (ib == 0 ? 1.0 : -1.0) * dl * bc;
  else if (c == 'D') factor = centered ? -2.0 * bc : -1.0 * bc;
  const int qqq = centered ? 0 : 1;
  const int i = (ib == 0) ? 1 : nhi;
  return factor * d2[i] / (d1[i - (1 - ib) + qqq] * d2[i]);   // src/solver.f90:125 (3-D weight)
}

static void setup_impdiff(Ctx &ctx) {
  const int nl = (int)ctx.boxes.size();
  auto fold_sign = [](char c, bool centered) {
    if (c == 'N') return 1.0;
    if (c == 'D') return centered ? -1.0 : 0.0;
    throw Error("a physical boundary must be 'D' or 'N'");
  };
  for (int c = 0; c < 3; ++c)
    for (int par = 0; par < 2; ++par) {
      std::vector<SBox> sb(nl);
      for (int lb = 0; lb < nl; ++lb) {
        Box &b = *ctx.boxes[lb];
        SBox &s = sb[lb];
        for (int d = 0; d < 3; ++d) {
          const bool cen = d != c;
          s.n[d] = b.g.n[d] - ((!cen && b.info.is_bound[d][1]) ? 1 : 0);   // hiu/hiv/hiw, src/main.f90:774-776
          s.d1[d] = cen ? b.h_dc[d] : b.h_df[d];
          s.d2[d] = cen ? b.h_df[d] : b.h_dc[d];
          s.qqq[d] = cen ? 0 : 1;
          for (int ib = 0; ib < 2; ++ib) {
            s.nb[d][ib] = b.info.nb[d][ib];
            s.phys[d][ib] = b.info.is_bound[d][ib] && ctx.prm.periods[d] == 0;
            s.sgn[d][ib] = s.phys[d][ib] ? fold_sign(b.info.cbcvel[c][d][ib], cen) : 0.0;
            if (par == 0)
              b.cbu[c][d][ib] = helm_bc_coef(b, d, ib, b.info.cbcvel[c][d][ib], b.info.bcvel[c][d][ib], cen, s.n[d], s.d1[d], s.d2[d],
                                             ctx.prm.periods);
          }
        }
        s.x = (c == 0 ? b.u[par] : (c == 1 ? b.v[par] : b.w[par])).p;
        s.b = b.rhs.p;
        s.sy = b.g.sy; s.sz = b.g.sz;
      }
      ctx.usolver[c][par].reset(new MGSolver);
      ctx.usolver[c][par]->setup(ctx, std::move(sb), false, 1, (c == 0 ? AK_U0 : (c == 1 ? AK_V0 : AK_W0)) + par);
    }
  for (int is = 0; is < ctx.prm.nscal; ++is) {
    std::vector<SBox> sb(nl);
    for (int lb = 0; lb < nl; ++lb) {
      Box &b = *ctx.boxes[lb];
      SBox &s = sb[lb];
      for (int d = 0; d < 3; ++d) {
        s.n[d] = b.g.n[d];
        s.d1[d] = b.h_dc[d];
        s.d2[d] = b.h_df[d];
        for (int ib = 0; ib < 2; ++ib) {
          s.nb[d][ib] = b.info.nb[d][ib];
          s.phys[d][ib] = b.info.is_bound[d][ib] && ctx.prm.periods[d] == 0;
          s.sgn[d][ib] = s.phys[d][ib] ? fold_sign(b.info.cbcscal[is][d][ib], true) : 0.0;
          b.cbs[is][d][ib] = helm_bc_coef(b, d, ib, b.info.cbcscal[is][d][ib], b.info.bcscal[is][d][ib], true, s.n[d], s.d1[d],
                                          s.d2[d], ctx.prm.periods);
        }
      }
      s.x = b.s[is].p;
      s.b = b.rhs.p;
      s.sy = b.g.sy; s.sz = b.g.sz;
    }
    ctx.ssolver[is].reset(new MGSolver);
    ctx.ssolver[is]->setup(ctx, std::move(sb), false, 1, AK_S0 + is);
  }
}

static void finish_setup(Ctx &ctx);

static void commit(Ctx &ctx) {
  if (ctx.committed) throw Error("snac_commit called twice");
  if (ctx.boxes.empty()) throw Error("snac_commit: no boxes were added");
  const int fa = ctx.prm.fft_axis - 1;
  const bool batched = ctx.prm.fft_axis > 0;
  const int nl = (int)ctx.boxes.size();
  std::vector<BlockInfo> blocks;
  if (batched) blocks = merge_blocks(ctx, assemble_blocks(ctx));
  ctx.nq = batched ? ctx.boxes[0]->g.n[fa] : 1;
  q_range(ctx, ctx.nq, ctx.comm.me, ctx.q0, ctx.q1);
  if (batched && ctx.q1 <= ctx.q0) throw Error("more processes than wavenumbers: nothing to solve on this process");
  if (ctx.comm.nproc > 1) {
    size_t bytes = 0;
    for (auto &bp : ctx.boxes) {
      bytes += (size_t)(8 + ctx.prm.nscal) * (bp->g.alloc_elems() * sizeof(double) + 256);
      bytes += (size_t)2 * (8 + ctx.prm.nscal) * ((size_t)(bp->g.n[1] + 2) * (bp->g.n[2] + 2) * sizeof(double) + 256);   // x-face landing buffers
    }
    if (!batched)   // 3-D solve: the iterate of every multigrid level lives in the arena (coarse levels sum to < 1 level 0)
      for (auto &bp : ctx.boxes) bytes += (size_t)(2.2 * (bp->g.alloc_elems() * sizeof(double))) + 32 * 256;
    for (BlockInfo &B : blocks) {
      int pd[2], q = 0;
      for (int d = 0; d < 3; ++d) if (d != fa) pd[q++] = d;
      Grid3 sg = Grid3::make(B.hi_g[pd[0]] - B.lo_g[pd[0]] + 1, B.hi_g[pd[1]] - B.lo_g[pd[1]] + 1, std::max(ctx.q1 - ctx.q0, 1));
      bytes += 5 * (sg.alloc_elems() * sizeof(double) + 256);   // right-hand side + the four arrays of MGSolver::xslot
    }
    ctx.comm.init(ctx.prm.nproc, ctx.prm.myproc, bytes);
  }
  for (auto &bp : ctx.boxes) {
    Box &b = *bp;
    const int r = b.info.rank;
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
    for (int q = 0; q < 2; ++q) {
      alloc_field(ctx, b.u[q], b.g, r, AK_U0 + q);
      alloc_field(ctx, b.v[q], b.g, r, AK_V0 + q);
      alloc_field(ctx, b.w[q], b.g, r, AK_W0 + q);
    }
    alloc_field(ctx, b.p, b.g, r, AK_P);
    b.rhs.alloc(b.g);
    b.pp_is_po = !batched;
    if (b.pp_is_po) alloc_field(ctx, b.po, b.g, r, AK_PO);
    else { alloc_field(ctx, b.pp, b.g, r, AK_PP); b.stage.alloc(b.g); }
    b.du.alloc(b.g); b.dv.alloc(b.g); b.dw.alloc(b.g);
    for (int s = 0; s < ctx.prm.nscal; ++s) { alloc_field(ctx, b.s[s], b.g, r, AK_S0 + s); b.ds[s].alloc(b.g); }
    if (ctx.prm.nscal > 0) { b.scr0.alloc(b.g); b.scr1.alloc(b.g); }
    compute_bc_rhs(ctx, b);
    b.tma_ok = true;
    for (int q = 0; q < 2 && b.tma_ok; ++q) {
      b.tma_ok = mom_tma_make_map(b.tmap[q][0], b.u[q].p, b.g) && mom_tma_make_map(b.tmap[q][1], b.v[q].p, b.g) &&
                 mom_tma_make_map(b.tmap[q][2], b.w[q].p, b.g);
    }
    b.tma_ok = b.tma_ok && mom_tma_make_map(b.tmap[0][3], b.p.p, b.g) && mom_tma_make_map(b.tmap_hist[0], b.du.p, b.g, true) &&
               mom_tma_make_map(b.tmap_hist[1], b.dv.p, b.g, true) && mom_tma_make_map(b.tmap_hist[2], b.dw.p, b.g, true);
    b.tma_pois_ok = b.tma_ok && pois_tma_setup(b);
    if (ctx.comm.nproc > 1) {   // landing buffers of the x-faces pushed by neighbours on other processes (bound.cu)
      std::vector<int> fkeys = {AK_U0, AK_U1, AK_V0, AK_V1, AK_W0, AK_W1, AK_P};
      fkeys.push_back(b.pp_is_po ? AK_PO : AK_PP);
      for (int s = 0; s < ctx.prm.nscal; ++s) fkeys.push_back(AK_S0 + s);
      for (int ib = 0; ib < 2; ++ib) {
        const int nbr = b.info.nb[0][ib];
        if (nbr < 0 || ctx.owner[nbr] == ctx.comm.me) continue;
        for (int k : fkeys) ctx.comm.alloc((size_t)(b.g.n[1] + 2) * (b.g.n[2] + 2), r, AK_XSTAGE + 2 * k + ib);
      }
    }
  }
  // transforms (fftini for the pressure, src/main.f90:735)
  if (batched) {
    Box &b0 = *ctx.boxes[0];
    const char bc[2] = {b0.info.cbcpre[fa][0], b0.info.cbcpre[fa][1]};
    for (int r = 0; r < ctx.prm.nrank; ++r) {
      const snac_box &w = ctx.world[r];
      if (w.hi[fa] - w.lo[fa] + 1 != b0.g.n[fa] || w.cbcpre[fa][0] != bc[0] || w.cbcpre[fa][1] != bc[1])
        throw Error("all boxes must have the same size and boundary conditions along the transform axis (src/sanity.f90:66-100)");
    }
    r2r_plan_init(ctx.plan_p, fa, b0.g.n[fa], bc, true);
  }
  {  // the pressure problem as the solver sees it
    MGSolver &S = ctx.psolver;
    std::vector<SBox> sb;
    if (batched) {
      // canonical layout: coupled directions first, the transform direction (batch) last; one solver
      // box per block; this process solves the wavenumbers q0..q1-1
      int q = 0;
      for (int d = 0; d < 3; ++d) if (d != fa) S.perm[q++] = d;
      S.perm[2] = fa;
      sb.resize(blocks.size());
      ctx.block_index.assign(1, -1);
      for (size_t ib_ = 0; ib_ < blocks.size(); ++ib_) {
        BlockInfo &B = blocks[ib_];
        for (int mid : B.members) {
          if ((int)ctx.block_index.size() <= mid) ctx.block_index.resize(mid + 1, -1);
          ctx.block_index[mid] = (int)ib_;
        }
        SBox &s = sb[ib_];
        const snac_box &w0 = ctx.world[B.face_rank[fa][0]];
        for (int sd = 0; sd < 3; ++sd) {
          const int d = S.perm[sd];
          s.n[sd] = sd == 2 ? (ctx.q1 - ctx.q0) : (B.hi_g[d] - B.lo_g[d] + 1);
          s.d1[sd] = B.dc[d];
          s.d2[sd] = B.df[d];
          for (int ib = 0; ib < 2; ++ib) {
            const snac_box &wf = ctx.world[B.face_rank[d][ib]];
            s.nb[sd][ib] = B.nb[d][ib];
            s.phys[sd][ib] = wf.is_bound[d][ib] && ctx.prm.periods[d] == 0;
            s.sgn[sd][ib] = 0.0;
            if (s.phys[sd][ib] && sd != 2) {   // src/solver.f90:71-94 (centred unknowns)
              const char c = wf.cbcpre[d][ib];
              if (c == 'N') s.sgn[sd][ib] = 1.0;
              else if (c == 'D') s.sgn[sd][ib] = -1.0;
              else throw Error("pressure boundary condition on a physical face must be 'D' or 'N'");
            }
          }
        }
        // eigenvalues / dl^2 of this process' wavenumbers (src/fft.f90:151-189, src/solver.f90:1124-1138)
        const char bc[2] = {w0.cbcpre[fa][0], w0.cbcpre[fa][1]};
        std::vector<double> l = r2r_eigenvalues(ctx.nq, bc, true);
        const double dl = w0.dl[fa][0];
        s.lam.assign((size_t)(ctx.q1 - ctx.q0) + 2, 0.0);
        for (int k = 0; k < ctx.q1 - ctx.q0; ++k) s.lam[k + 1] = l[ctx.q0 + k * q_stride(ctx)] / (dl * dl);
        if (ctx.comm.nproc > 1) {   // slabs live in the arena: peers write the right-hand side and read the solution
          Grid3 sg = Grid3::make(s.n[0], s.n[1], std::max(s.n[2], 1));
          s.x = ctx.comm.alloc(sg.alloc_elems(), B.id, AK_SLAB_X) + 15;
          s.b = ctx.comm.alloc(sg.alloc_elems(), B.id, AK_SLAB_B) + 15;
          s.xalt[0] = ctx.comm.alloc(sg.alloc_elems(), B.id, AK_SLAB_X1) + 15;
          s.xalt[1] = ctx.comm.alloc(sg.alloc_elems(), B.id, AK_SLAB_K1) + 15;
          s.xalt[2] = ctx.comm.alloc(sg.alloc_elems(), B.id, AK_SLAB_K2) + 15;
        }
      }
      ctx.blocks_lo.clear();
      ctx.blocks_id.clear();
      for (BlockInfo &B : blocks) { ctx.blocks_lo.push_back({B.lo_g[0], B.lo_g[1], B.lo_g[2]}); ctx.blocks_id.push_back(B.id); }
    } else {
      sb.resize(nl);
      for (int lb = 0; lb < nl; ++lb) {
        Box &b = *ctx.boxes[lb];
        SBox &s = sb[lb];
        for (int d = 0; d < 3; ++d) {
          s.n[d] = b.g.n[d];
          s.d1[d] = b.h_dc[d];
          s.d2[d] = b.h_df[d];
          for (int ib = 0; ib < 2; ++ib) {
            s.nb[d][ib] = b.info.nb[d][ib];
            s.phys[d][ib] = b.info.is_bound[d][ib] && ctx.prm.periods[d] == 0;
            s.sgn[d][ib] = 0.0;
            if (s.phys[d][ib]) {
              const char c = b.info.cbcpre[d][ib];
              if (c == 'N') s.sgn[d][ib] = 1.0;
              else if (c == 'D') s.sgn[d][ib] = -1.0;
              else throw Error("pressure boundary condition on a physical face must be 'D' or 'N'");
            }
          }
        }
        s.x = b.po.p;
        s.b = b.rhs.p;
      }
    }
    // separable planes (one rectangle per plane with a uniform, transformable direction): the direct solver
    // replaces the multigrid cycle; the hierarchy is then not built (level 0 only holds the arrays)
    int tdir = -1;
    if (batched && !ctx.prm.mg_generic && ctx.prm.plane_solver != 1) tdir = direct_pick(sb, true, ctx.local_of_rank);
    if (ctx.prm.plane_solver == 2 && tdir < 0)
      throw Error("plane_solver = direct: the 2-D problems are not separable (planes split over blocks, or no uniform "
                  "periodic/Neumann/Dirichlet direction); use the multigrid");
    S.setup(ctx, std::move(sb), batched, tdir >= 0 ? 1 : 24, batched ? -1 : AK_PO);
    if (tdir >= 0) { S.direct.setup(ctx, S, tdir); ctx.direct_on = true; }
  }
  if (ctx.prm.impdiff) setup_impdiff(ctx);
  ctx.red.alloc(8);
  SNAC_CUDA(cudaMallocHost(&ctx.h_red, 4 * MAXPROC * sizeof(double)));
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
  ctx.committed = true;
  if (ctx.comm.nproc == 1) finish_setup(ctx);
}

// halo tables, boundary-condition tables and transform routing: needs the peers' arenas when
// several processes take part (called from snac_comm_connect in that case)
static void finish_setup(Ctx &ctx) {
  const int nl = (int)ctx.boxes.size();
  const int fa = ctx.prm.fft_axis - 1;
  ctx.psolver.connect(ctx);
  for (int c = 0; c < 3; ++c)
    for (int par = 0; par < 2; ++par)
      if (ctx.usolver[c][par]) ctx.usolver[c][par]->connect(ctx);
  for (int is = 0; is < SNAC_MAX_SCAL; ++is)
    if (ctx.ssolver[is]) ctx.ssolver[is]->connect(ctx);
  for (int par = 0; par < 2; ++par) {
    std::vector<ExField> ef;
    for (int f = 0; f < 3; ++f)
      for (int lb = 0; lb < nl; ++lb) {
        Box &b = *ctx.boxes[lb];
        ef.push_back(ExField{lb, f == 0 ? b.u[par].p : (f == 1 ? b.v[par].p : b.w[par].p), b.g});
      }
    const std::vector<int> keys = {AK_U0 + par, AK_V0 + par, AK_W0 + par};
    ctx.ex_uvw[par].build(ctx, ef, nullptr, &keys);
    // bounduvw(.true.) and boundp(p) follow each other at the end of a sub-step (src/main.f90:1025-1032) and
    // touch different arrays: one exchange for the four of them (a third fewer launches and handshakes)
    for (int lb = 0; lb < nl; ++lb) ef.push_back(ExField{lb, ctx.boxes[lb]->p.p, ctx.boxes[lb]->g});
    const std::vector<int> keys4 = {AK_U0 + par, AK_V0 + par, AK_W0 + par, AK_P};
    ctx.ex_uvwp[par].build(ctx, ef, nullptr, &keys4);
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
    const std::vector<int> kp = {AK_P}, kpp = {ctx.boxes[0]->pp_is_po ? AK_PO : AK_PP};
    ctx.ex_p.build(ctx, ep, nullptr, &kp);
    ctx.ex_pp.build(ctx, epp, nullptr, &kpp);
    auto cbc = [](Box &b, int, int d, int ib) -> char { return b.info.cbcpre[d][ib]; };
    auto bc = [](Box &b, int, int d, int ib) -> double { return b.info.bcpre[d][ib]; };
    build_cen_bc(ctx, ctx.bc_p, [](Box &b, int) { return b.p.p; }, 0, cbc, bc);
    build_cen_bc(ctx, ctx.bc_pp, [](Box &b, int) { return b.pp_ptr(); }, 0, cbc, bc);
  }
  for (int s = 0; s < ctx.prm.nscal; ++s) {
    std::vector<ExField> es;
    for (int lb = 0; lb < nl; ++lb) es.push_back(ExField{lb, ctx.boxes[lb]->s[s].p, ctx.boxes[lb]->g});
    const std::vector<int> ks = {AK_S0 + s};
    ctx.ex_s[s].build(ctx, es, nullptr, &ks);
    auto cbc = [](Box &b, int s_, int d, int ib) -> char { return b.info.cbcscal[s_][d][ib]; };
    auto bc = [](Box &b, int s_, int d, int ib) -> double { return b.info.bcscal[s_][d][ib]; };
    build_cen_bc(ctx, ctx.bc_s[s], [](Box &b, int s_) { return b.s[s_].p; }, s, cbc, bc);
  }
  // transform routing: where each local box's pencils go to / come from, per owning process
  if (ctx.prm.fft_axis > 0) {
    MGSolver &S = ctx.psolver;
    ctx.route.assign(nl, FFTRoute());
    for (int lb = 0; lb < nl; ++lb) {
      Box &b = *ctx.boxes[lb];
      const int bi = ctx.block_index[b.info.block];
      const Grid3 &sg = S.grid0(bi);
      FFTRoute &R = ctx.route[lb];
      for (int sd = 0; sd < 3; ++sd) R.st[S.perm[sd]] = sg.st(sd);
      // offset of this box's first cell inside the block's array (coupled directions only)
      long off = 1 + sg.sy + sg.sz;
      for (int sd = 0; sd < 2; ++sd) off += (long)(b.info.lo[S.perm[sd]] - ctx.blocks_lo[bi][S.perm[sd]]) * sg.st(sd);
      R.chunk = (ctx.nq + ctx.comm.nproc - 1) / ctx.comm.nproc;
      R.off = off;
      R.nsplit = ctx.comm.nproc;
      R.cyclic = q_is_cyclic(ctx);
      for (int p = 0; p < ctx.comm.nproc; ++p) {
        int q0, q1;
        q_range(ctx, ctx.nq, p, q0, q1);
        double *xb, *bb;
        if (p == ctx.comm.me) { xb = S.xcur(bi); bb = S.bcur(bi); }
        else { xb = ctx.comm.remote(p, ctx.blocks_id[bi], AK_SLAB_X) + 15; bb = ctx.comm.remote(p, ctx.blocks_id[bi], AK_SLAB_B) + 15; }
        // every array the owner's iterate can move to (the owner's own ones are looked up when they are needed)
        static const int slot_keys[4] = {AK_SLAB_X, AK_SLAB_X1, AK_SLAB_K1, AK_SLAB_K2};
        for (int sl = 0; sl < 4; ++sl)
          R.xs[p][sl] = (p == ctx.comm.me || ctx.comm.nproc == 1) ? nullptr
                        : ctx.comm.remote(p, ctx.blocks_id[bi], slot_keys[sl]) + 15 + off - (q_is_cyclic(ctx) ? 0 : (long)q0 * sg.st(2));
        // chunks: the wavenumber index is absolute in the kernels, shift by the slab origin; cyclic: the
        // kernels divide the index by the number of processes themselves
        const long shift = q_is_cyclic(ctx) ? 0 : (long)q0 * sg.st(2);
        R.x[p] = xb + off - shift;
        R.b[p] = bb + off - shift;
      }
    }
    (void)fa;
  }
  ctx.ready = true;
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

__global__ void permute_copy_kernel(double *dst, long d0, long d1, long d2, const double *src, long s0, long s1, long s2,
                                    int n0, int n1, int n2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y, k = blockIdx.z;
  if (i < n0 && j < n1 && k < n2) dst[i * d0 + j * d1 + k * d2] = src[i * s0 + j * s1 + k * s2];
}

static void transfer(Ctx &ctx, int rank, int field, double *host, bool up) {
  if (up) ctx.vel_epoch++;
  Box &b = ctx.box_of_rank(rank);
  const int lb = ctx.local_of_rank[rank];
  const bool batched = ctx.prm.fft_axis > 0;
  // spectral arrays live in the solver's canonical per-block layout: go through a staging array
  const bool spectral = batched && (field == SNAC_FIELD_PO || field == SNAC_FIELD_RHS_HAT);
  if (spectral && ctx.comm.nproc > 1) throw Error("spectral fields are distributed over the processes: transfer is single-process only");
  Field3 *f;
  bool dummy;
  if (spectral) f = &b.stage;
  else if (field == SNAC_FIELD_RHS_HAT) f = &b.rhs;
  else f = field_of(ctx, b, field, dummy);
  const long pint = 1 + b.g.sy + b.g.sz;
  dim3 blk(32, 8, 1), grd((b.g.n[0] + 31) / 32, (b.g.n[1] + 7) / 8, b.g.n[2]);
  double *sp = nullptr;
  const long *sst = nullptr;
  if (spectral) {
    const FFTRoute &R = ctx.route[lb];
    sst = R.st;
    sp = field == SNAC_FIELD_PO ? R.x[0] : R.b[0];   // first interior element of this box inside its block's array
    if (!up) {
      SNAC_CUDA(cudaMemsetAsync(b.stage.buf.raw, 0, b.stage.buf.elems * sizeof(double), ctx.stream));
      permute_copy_kernel<<<grd, blk, 0, ctx.stream>>>(b.stage.p + pint, 1, b.g.sy, b.g.sz, sp, sst[0], sst[1], sst[2],
                                                       b.g.n[0], b.g.n[1], b.g.n[2]);
    }
  }
  // host <-> device through a dense staging buffer: one contiguous PCIe copy at full rate plus a
  // re-pitching kernel (a pitched 3-D copy is split into one DMA descriptor per 2 KB row and runs
  // at about half the PCIe bandwidth)
  const int w = b.g.n[0] + 2, hh = b.g.n[1] + 2, dpt = b.g.n[2] + 2;
  const size_t dense = (size_t)w * hh * dpt;
  if (ctx.xfer.elems < dense) ctx.xfer.alloc(dense);
  dim3 gfull((w + 31) / 32, (hh + 7) / 8, dpt);
  if (up) {
    SNAC_CUDA(cudaMemcpyAsync(ctx.xfer.raw, host, dense * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    permute_copy_kernel<<<gfull, blk, 0, ctx.stream>>>(f->p, 1, b.g.sy, b.g.sz, ctx.xfer.raw, 1, w, (long)w * hh, w, hh, dpt);
  } else {
    permute_copy_kernel<<<gfull, blk, 0, ctx.stream>>>(ctx.xfer.raw, 1, w, (long)w * hh, f->p, 1, b.g.sy, b.g.sz, w, hh, dpt);
    SNAC_CUDA(cudaMemcpyAsync(host, ctx.xfer.raw, dense * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  }
  if (spectral && up)
    permute_copy_kernel<<<grd, blk, 0, ctx.stream>>>(sp, sst[0], sst[1], sst[2], b.stage.p + pint, 1, b.g.sy, b.g.sz,
                                                     b.g.n[0], b.g.n[1], b.g.n[2]);
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
}

// Outputs and checkpoints without stopping the time loop (the reference writes
// them with blocking MPI-IO from the host arrays, src/load.f90:84-187, src/output.f90:227-257): what the field holds
// NOW is packed into a dense staging array in stream order, the device -> host copy runs on a copy stream, and the
// caller gets a ticket to wait on -- from any host thread, it touches nothing but the ticket's event.
static int download_async(Ctx &ctx, int rank, int field, double *host) {
  Box &b = ctx.box_of_rank(rank);
  if (ctx.prm.fft_axis > 0 && (field == SNAC_FIELD_PO || field == SNAC_FIELD_RHS_HAT))
    throw Error("download_async: spectral fields are not supported (use snac_download)");
  bool dummy;
  Field3 *f = field == SNAC_FIELD_RHS_HAT ? &b.rhs : field_of(ctx, b, field, dummy);
  if (!ctx.copy_stream) SNAC_CUDA(cudaStreamCreateWithFlags(&ctx.copy_stream, cudaStreamNonBlocking));
  int slot = -1;
  for (int q = 0; q < Ctx::MAX_ASYNC && slot < 0; ++q) {
    Ctx::AsyncSlot &S = ctx.async_slot[q];
    if (!S.done) {
      SNAC_CUDA(cudaEventCreateWithFlags(&S.packed, cudaEventDisableTiming));
      SNAC_CUDA(cudaEventCreateWithFlags(&S.done, cudaEventDisableTiming));
      slot = q;
    } else if (cudaEventQuery(S.done) == cudaSuccess) {
      slot = q;
    }
  }
  (void)cudaGetLastError();   // cudaErrorNotReady of the queries above is not an error
  if (slot < 0) throw Error("download_async: too many copies in flight (wait for some of them)");
  Ctx::AsyncSlot &S = ctx.async_slot[slot];
  const int w = b.g.n[0] + 2, hh = b.g.n[1] + 2, dpt = b.g.n[2] + 2;
  const size_t dense = (size_t)w * hh * dpt;
  if (S.stage.elems < dense) S.stage.alloc(dense);
  dim3 blk(32, 8, 1), gfull((w + 31) / 32, (hh + 7) / 8, dpt);
  permute_copy_kernel<<<gfull, blk, 0, ctx.stream>>>(S.stage.raw, 1, w, (long)w * hh, f->p, 1, b.g.sy, b.g.sz, w, hh, dpt);
  ctx.launches++;
  SNAC_CUDA(cudaEventRecord(S.packed, ctx.stream));
  SNAC_CUDA(cudaStreamWaitEvent(ctx.copy_stream, S.packed, 0));
  SNAC_CUDA(cudaMemcpyAsync(host, S.stage.raw, dense * sizeof(double), cudaMemcpyDeviceToHost, ctx.copy_stream));
  SNAC_CUDA(cudaEventRecord(S.done, ctx.copy_stream));
  return slot;
}

// ---- pieces of the sub-step ------------------------------------------------------------------
static void do_rk_mom(Ctx &ctx, int irk, double dt) {
  ctx.vel_epoch++;
  if (irk < 1 || irk > 3) throw Error("rk_mom: irk must be 1, 2 or 3");
  double pts = 0.0;
  for (auto &b : ctx.boxes) pts += (double)b->g.n[0] * b->g.n[1] * b->g.n[2];
  ProfScope ps(ctx, PC_RK_MOM, pts);
  for (auto &b : ctx.boxes)
    if (!launch_rk_mom_tma(ctx, *b, irk, dt)) launch_rk_mom(ctx, *b, irk, dt);
  ctx.par = 1 - ctx.par;
}
static void do_bounduvw(Ctx &ctx, bool is_correc) {
  ctx.vel_epoch++;
  ProfScope ps(ctx, PC_HALO_BC);
  ctx.ex_uvw[ctx.par].run_all(ctx);
  ctx.bc_uvw[ctx.par][is_correc ? 1 : 0].run(ctx);
}
// inflow (src/bound.f90:438-477): prescribed velocity planes on the 'D' faces flagged by
// is_bound_inflow (src/main.f90:594-632); the normal component only when not correcting
static void build_inflow(Ctx &ctx) {
  for (int par = 0; par < 2; ++par)
    for (int ic = 0; ic < 2; ++ic) {
      std::vector<FaceOp> ops[3];
      for (auto &bp : ctx.boxes) {
        Box &b = *bp;
        double *fld[3] = {b.u[par].p, b.v[par].p, b.w[par].p};
        for (int idir = 0; idir < 3; ++idir)
          for (int ib = 0; ib < 2; ++ib) {
            if (!b.is_inflow[idir][ib]) continue;
            const int a = idir == 0 ? 1 : 0, bb = idir == 2 ? 1 : 2;
            for (int c = 0; c < 3; ++c) {
              const bool normal = (c == idir);
              if (normal && ic) continue;
              FaceOp op;
              op.p = fld[c];
              op.sd = b.g.st(idir); op.sa = b.g.st(a); op.sb = b.g.st(bb);
              op.n = b.g.n[idir]; op.na = b.g.n[a] + 2; op.nb = b.g.n[bb] + 2;
              op.factor = 0.0; op.sgn = 0.0;
              op.plane = b.inflow[idir][ib][c].p();
              op.kind = normal ? (ib == 0 ? FACE_INFLOW_N_LO : FACE_INFLOW_N_HI) : (ib == 0 ? FACE_INFLOW_T_LO : FACE_INFLOW_T_HI);
              ops[idir].push_back(op);
            }
          }
      }
      ctx.bc_inflow[par][ic].upload(ops);
    }
  ctx.inflow_dirty = false;
}
static void do_inflow(Ctx &ctx, bool is_correc) {
  if (!ctx.has_inflow) return;
  ctx.vel_epoch++;
  ProfScope ps(ctx, PC_HALO_BC);
  if (ctx.inflow_dirty) build_inflow(ctx);
  ctx.bc_inflow[ctx.par][is_correc ? 1 : 0].run(ctx);
}
static void set_inflow(Ctx &ctx, int rank, int idir, int ib, const double *veln, const double *velt1, const double *velt2) {
  if (idir < 0 || idir > 2 || ib < 0 || ib > 1) throw Error("set_inflow: idir must be 0..2 and ib 0 or 1");
  if (ctx.prm.impdiff) throw Error("set_inflow: inflow planes with implicit diffusion are not supported");
  Box &b = ctx.box_of_rank(rank);
  if (!b.info.is_bound[idir][ib] || b.info.cbcvel[idir][idir][ib] != 'D')
    throw Error("set_inflow: the face is not a physical boundary with a Dirichlet normal velocity (src/main.f90:596)");
  const int a = idir == 0 ? 1 : 0, bb = idir == 2 ? 1 : 2;
  const size_t n = (size_t)(b.g.n[a] + 2) * (b.g.n[bb] + 2);
  // veln is the component normal to the face, velt1/velt2 the other two in x<y<z order (src/main.f90:601-621)
  const double *src[3];
  src[idir] = veln; src[a] = velt1; src[bb] = velt2;
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
  for (int c = 0; c < 3; ++c) b.inflow[idir][ib][c].upload(std::vector<double>(src[c], src[c] + n));
  b.is_inflow[idir][ib] = true;
  ctx.has_inflow = true;
  ctx.inflow_dirty = true;
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
  if (!ctx.ready) throw Error("snac_comm_connect has not been called");
  double pts = 0.0;
  for (auto &b : ctx.boxes) pts += (double)b->g.n[0] * b->g.n[1] * b->g.n[2];
  ProfScope ps(ctx, dir > 0 ? PC_FFT_FWD : PC_FFT_BWD, pts);
  for (int lb = 0; lb < (int)ctx.boxes.size(); ++lb) {
    Box &b = *ctx.boxes[lb];
    const FFTRoute &R = ctx.route[lb];
    const long pst[3] = {1, b.g.sy, b.g.sz};
    const long pint = 1 + b.g.sy + b.g.sz;   // first interior element
    if (dir > 0) {        // physical right-hand side -> the owners' slabs
      const double *in[1] = {b.rhs.p + pint};
      r2r_execute(ctx, ctx.plan_p, +1, b.g.n, in, 1, pst, R.b, R.nsplit, R.st, R.chunk, 1.0, R.cyclic);
    } else {              // the owners' solution slabs -> physical pp, times normfft (src/main.f90:1017-1020)
      double *out[1] = {b.pp.p + pint};
      r2r_execute(ctx, ctx.plan_p, -1, b.g.n, R.x, R.nsplit, R.st, out, 1, pst, R.chunk, ctx.plan_p.normfft, R.cyclic);
    }
  }
  // forward: every process must have received all pieces of its slabs before solving
  if (dir > 0) ctx.comm.handshake(ctx, CH_FFT, ~0u);
}
static void do_solve(Ctx &ctx) {
  ctx.psolver.solve(ctx);
  if (ctx.psolver.swapped) {   // the solution lives in another of the solver's arrays now: here and, alike, on every peer
    MGSolver &S = ctx.psolver;
    for (int lb = 0; lb < (int)ctx.boxes.size(); ++lb) {
      FFTRoute &R = ctx.route[lb];
      const int bi = ctx.block_index[ctx.boxes[lb]->info.block];
      for (int p = 0; p < std::max(R.nsplit, 1); ++p) {
        if (p == ctx.comm.me || ctx.comm.nproc == 1) {
          int q0, q1;
          q_range(ctx, ctx.nq, ctx.comm.me, q0, q1);
          R.x[p] = S.xcur(bi) + R.off - (q_is_cyclic(ctx) || ctx.comm.nproc == 1 ? 0 : (long)q0 * S.grid0(bi).st(2));
        } else {
          R.x[p] = R.xs[p][S.slot_primary];
        }
      }
    }
    S.swapped = false;
  }
  // batched: the peers read this process' solution slab in the backward transform
  if (ctx.prm.fft_axis > 0) ctx.comm.handshake(ctx, CH_SOLVED, ~0u);
}

// after a stream synchronisation: did a handshake give up on a peer (comm.cu)?
static void check_peers(Ctx &ctx) {
  const int e = ctx.comm.error_state();
  if (e == 1) throw Error("a peer process reported an error and left the run (see its snac_last_error)");
  if (e == 2) throw Error("a peer process did not arrive at a handshake within the time limit (dead or hung)");
}

__global__ void publish_kernel(const double *vals, double *const *slots, int n, int nproc) {
  const int p = threadIdx.x;
  if (p < nproc) for (int i = 0; i < n; ++i) slots[p][i] = vals[i];
  __threadfence_system();   // performed at the peers before the kernel completes (the flag is written by a later kernel)
}

// MPI_ALLREDUCE of up to 3 values that sit in device memory: every process writes them into its slot of every peer's
// reduction table (double-buffered by the parity of the sequence number), one handshake, and the table is read back
void Ctx::peer_reduce(const double *d_vals, int n, double *sum, double *mx) {
  if (n < 1 || n > 3) throw Error("peer_reduce: 1 to 3 values");
  const int np = comm.nproc;
  if (np == 1) {
    SNAC_CUDA(cudaMemcpyAsync(h_red, d_vals, n * sizeof(double), cudaMemcpyDeviceToHost, stream));
    SNAC_CUDA(cudaStreamSynchronize(stream));
    for (int q = 0; q < n; ++q) { sum[q] = h_red[q]; mx[q] = h_red[q]; }
    return;
  }
  if (!ready) throw Error("snac_comm_connect has not been called");
  const size_t base = PeerComm::FLAGS_BYTES + (size_t)((comm.seq[CH_REDUCE] + 1) & 1) * MAXPROC * 4 * sizeof(double);
  std::vector<double *> slots(np);
  for (int p = 0; p < np; ++p) slots[p] = reinterpret_cast<double *>(comm.peer[p] + base) + 4 * comm.me;
  SNAC_CUDA(cudaMemcpyAsync(d_slots, slots.data(), np * sizeof(double *), cudaMemcpyHostToDevice, stream));
  publish_kernel<<<1, 32, 0, stream>>>(d_vals, d_slots, n, np);
  launches++;
  comm.handshake(*this, CH_REDUCE, ~0u);
  SNAC_CUDA(cudaMemcpyAsync(h_red, comm.arena + base, 4 * MAXPROC * sizeof(double), cudaMemcpyDeviceToHost, stream));
  SNAC_CUDA(cudaStreamSynchronize(stream));
  check_peers(*this);
  for (int q = 0; q < n; ++q) { sum[q] = 0.0; mx[q] = -1e300; }
  for (int p = 0; p < np; ++p)
    for (int q = 0; q < n; ++q) { sum[q] += h_red[4 * p + q]; mx[q] = std::max(mx[q], h_red[4 * p + q]); }
}

// local reductions, then the MPI_ALLREDUCEs of chkdt/chkdiv (src/chkdt.f90:48, src/chkdiv.f90:42-43)
static void do_chk(Ctx &ctx, bool want_dt, double *out3) {
  if (ctx.chk_epoch == ctx.vel_epoch) {   // (the same decision on every process: they make the same calls)
    for (int q = 0; q < 3; ++q) out3[q] = ctx.chk_out[q];
    return;
  }
  SNAC_CUDA(cudaMemsetAsync(ctx.red.raw, 0, 8 * sizeof(double), ctx.stream));
  {
    ProfScope ps(ctx, PC_CHK);
    for (auto &b : ctx.boxes) launch_chk(ctx, *b, ctx.red.raw);
  }
  (void)want_dt;
  double sum[3], mx[3];
  ctx.peer_reduce(ctx.red.raw, 3, sum, mx);
  out3[0] = mx[0]; out3[1] = sum[1]; out3[2] = mx[2];
  for (int q = 0; q < 3; ++q) ctx.chk_out[q] = out3[q];
  ctx.chk_epoch = ctx.vel_epoch;
}

static const double RKC[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};

// rk_scal + bulk_forcing_s for one scalar (src/main.f90:910-913)
static void do_rk_scal(Ctx &ctx, int irk, double dt, int is) {
  if (is < 0 || is >= ctx.prm.nscal) throw Error("rk_scal: scalar index out of range");
  if (irk < 1 || irk > 3) throw Error("rk_scal: irk must be 1, 2 or 3");
  const double f1 = RKC[irk - 1][0] * dt, f2 = RKC[irk - 1][1] * dt, f12 = f1 + f2;
  const bool forced = ctx.prm.is_sforced[is] != 0;
  for (auto &b : ctx.boxes) launch_scal_rhs(ctx, *b, is, 1.0 / ctx.prm.alphai[is]);
  double *d_sum = ctx.red.raw + 4;
  if (forced) SNAC_CUDA(cudaMemsetAsync(d_sum, 0, sizeof(double), ctx.stream));
  for (auto &b : ctx.boxes) launch_scal_update(ctx, *b, is, f1, f2, f12, d_sum);
  double f = 0.0;
  if (forced) {   // bulk_mean_s (src/scal.f90:156-179; MPI_ALLREDUCE :177)
    double sum[3], mx[3];
    ctx.peer_reduce(d_sum, 1, sum, mx);
    f = ctx.prm.scalf[is] - sum[0] / ctx.prm.vol_all;
  }
  if (ctx.prm.impdiff || forced)
    for (auto &b : ctx.boxes) launch_scal_finish(ctx, *b, is, f12, f);
  ctx.fs[is] += f;
}

// solve_impdiff_field_3d (src/solve_helmholtz.f90:102-138) for a scalar or a velocity component
static void do_helmholtz(Ctx &ctx, MGSolver &S, int comp /* 0..2 velocity, -1-is scalar */, double alphai) {
  const int nl = (int)ctx.boxes.size();
  for (int lb = 0; lb < nl; ++lb) {
    Box &b = *ctx.boxes[lb];
    const double *wq[3];
    int nhi[3];
    for (int d = 0; d < 3; ++d) {
      wq[d] = (comp >= 0 && d == comp) ? b.dc[d].p() : b.df[d].p();
      nhi[d] = S.sboxes[lb].n[d];
    }
    const double *fld = S.levels[0]->boxes[lb].xp;
    launch_helm_rhs(ctx, b, fld, wq, comp >= 0 ? b.cbu[comp] : b.cbs[-1 - comp], nhi, alphai);
  }
  S.alpha = alphai;
  S.solve(ctx);
}

static void substep(Ctx &ctx, int irk, double dt) {
  ctx.vel_epoch++;
  const double dtrk = (RKC[irk - 1][0] + RKC[irk - 1][1]) * dt;
  const double alpha = -ctx.prm.visc * dtrk / 2.0;
  if (!ctx.ready) throw Error("snac_comm_connect has not been called");
  for (int is = 0; is < ctx.prm.nscal; ++is) {   // src/main.f90:908-937
    do_rk_scal(ctx, irk, dt, is);
    if (ctx.prm.impdiff) do_helmholtz(ctx, *ctx.ssolver[is], -1 - is, 1.0 / (-(1.0 / ctx.prm.alphai[is]) * dtrk / 2.0));
    do_boundp(ctx, SNAC_FIELD_S0 + is);
  }
  do_rk_mom(ctx, irk, dt);
  if (ctx.prm.impdiff)   // src/main.f90:947-994
    for (int c = 0; c < 3; ++c) do_helmholtz(ctx, *ctx.usolver[c][ctx.par], c, 1.0 / alpha);
  do_bounduvw(ctx, false);
  do_inflow(ctx, false);   // src/main.f90:999
  double pts = 0.0;
  for (auto &b : ctx.boxes) pts += (double)b->g.n[0] * b->g.n[1] * b->g.n[2];
  {
    ProfScope ps(ctx, PC_FILLPS, pts);
    for (auto &b : ctx.boxes) launch_fillps_bc(ctx, *b, dtrk);
  }
  ctx.psolver.hist_key = irk - 1;   // the three pressure solves of a step behave differently, each like its predecessor a step earlier
  if (ctx.prm.fft_axis > 0) {
    do_fft(ctx, +1);
    do_solve(ctx);
    do_fft(ctx, -1);
  } else {
    do_solve(ctx);
  }
  do_boundp(ctx, SNAC_FIELD_PP);
  {
    ProfScope ps(ctx, PC_CORREC, pts);
    for (auto &b : ctx.boxes) launch_correc_updt(ctx, *b, dtrk, alpha, true, true);
  }
  {   // bounduvw(.true.), inflow(.true.), boundp(p): the halos of u,v,w,p in one exchange, then the face fills
    {
      ProfScope ps(ctx, PC_HALO_BC);
      ctx.ex_uvwp[ctx.par].run_all(ctx);
      ctx.bc_uvw[ctx.par][1].run(ctx);
    }
    do_inflow(ctx, true);    // src/main.f90:1028
    ProfScope ps(ctx, PC_HALO_BC);
    ctx.bc_p.run(ctx);
  }
}

}  // namespace snac

using namespace snac;

// every entry point ends with a look at the launch status: a refused launch (bad configuration) is not sticky and
// would otherwise surface at some later, unrelated check -- or never
#define SNAC_TRY(h, ...)                          \
  if (!(h)) return -2;                            \
  try {                                           \
    __VA_ARGS__;                                  \
    SNAC_CUDA(cudaGetLastError());                \
    return 0;                                     \
  } catch (const std::exception &e) {             \
    (h)->ctx.err = e.what();                      \
    (h)->ctx.comm.signal_abort();                 \
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
    ctx.world.assign(params->nrank, snac_box());
    ctx.world_set.assign(params->nrank, 0);
    for (int d = 0; d < 3; ++d) { ctx.world_dc[d].assign(params->nrank, {}); ctx.world_df[d].assign(params->nrank, {}); }
    ctx.comm.nproc = std::max(params->nproc, 1);
    ctx.comm.me = params->myproc;
    if (ctx.comm.nproc > MAXPROC) throw Error("snac_create: too many processes");
    SNAC_CUDA(cudaMalloc(&ctx.d_slots, MAXPROC * sizeof(double *)));
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
    if (ctx.world_set[box->rank]) throw Error("snac_add_box: box added twice");
    int n[3];
    for (int d = 0; d < 3; ++d) {
      n[d] = box->hi[d] - box->lo[d] + 1;
      if (n[d] < 1) throw Error("snac_add_box: empty box");
    }
    const double *dc[3] = {box->dxc, box->dyc, box->dzc}, *df[3] = {box->dxf, box->dyf, box->dzf};
    for (int d = 0; d < 3; ++d)
      if (!dc[d] || !df[d]) throw Error("snac_add_box: missing grid arrays");
    // every box of the run is recorded (several processes: neighbours and whole blocks must be known)
    ctx.world[box->rank] = *box;
    ctx.world[box->rank].dxc = ctx.world[box->rank].dxf = ctx.world[box->rank].dyc = nullptr;
    ctx.world[box->rank].dyf = ctx.world[box->rank].dzc = ctx.world[box->rank].dzf = nullptr;
    for (int d = 0; d < 3; ++d) {
      ctx.world_dc[d][box->rank].assign(dc[d], dc[d] + n[d] + 2);
      ctx.world_df[d][box->rank].assign(df[d], df[d] + n[d] + 2);
    }
    ctx.world_set[box->rank] = 1;
    if (ctx.owner[box->rank] != ctx.prm.myproc) return 0;   // owned by another process: description only
    std::unique_ptr<Box> b(new Box);
    b->info = ctx.world[box->rank];
    b->g = Grid3::make(n[0], n[1], n[2]);
    for (int d = 0; d < 3; ++d) {
      b->h_dc[d] = ctx.world_dc[d][box->rank];
      b->h_df[d] = ctx.world_df[d][box->rank];
    }
    ctx.local_of_rank[box->rank] = (int)ctx.boxes.size();
    ctx.boxes.push_back(std::move(b));
  })
}

int snac_commit(snac_handle *h) { SNAC_TRY(h, commit(h->ctx)) }

size_t snac_comm_blob_size(void) { return sizeof(CommBlob); }
int snac_comm_export(snac_handle *h, void *blob) {
  SNAC_TRY(h, {
    if (!h->ctx.committed) throw Error("snac_comm_export before snac_commit");
    if (h->ctx.comm.nproc == 1) { std::memset(blob, 0, sizeof(CommBlob)); return 0; }
    h->ctx.comm.export_blob(*static_cast<CommBlob *>(blob));
  })
}
int snac_comm_connect(snac_handle *h, const void *blobs, int nblobs) {
  SNAC_TRY(h, {
    Ctx &ctx = h->ctx;
    if (!ctx.committed) throw Error("snac_comm_connect before snac_commit");
    if (ctx.comm.nproc == 1) return 0;
    if (ctx.ready) throw Error("snac_comm_connect called twice");
    ctx.comm.connect(static_cast<const CommBlob *>(blobs), nblobs);
    finish_setup(ctx);
  })
}

int snac_finalize(snac_handle *h) {
  if (!h) return -2;
  cudaStreamSynchronize(h->ctx.stream);
  if (h->ctx.copy_stream) { cudaStreamSynchronize(h->ctx.copy_stream); cudaStreamDestroy(h->ctx.copy_stream); }
  for (auto &S : h->ctx.async_slot) {
    if (S.packed) cudaEventDestroy(S.packed);
    if (S.done) cudaEventDestroy(S.done);
  }
  if (h->ctx.h_red) cudaFreeHost(h->ctx.h_red);
  if (h->ctx.d_slots) cudaFree(h->ctx.d_slots);
  h->ctx.comm.release();
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

int snac_download_async(snac_handle *h, int rank, int field, double *host, int *ticket) {
  SNAC_TRY(h, *ticket = download_async(h->ctx, rank, field, host))
}
int snac_download_wait(snac_handle *h, int ticket) {
  if (!h || ticket < 0 || ticket >= Ctx::MAX_ASYNC || !h->ctx.async_slot[ticket].done) return -2;
  return cudaEventSynchronize(h->ctx.async_slot[ticket].done) == cudaSuccess ? 0 : -1;
}
void *snac_host_alloc(size_t bytes) {
  void *p = nullptr;
  return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr;
}
void snac_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int snac_rk_mom(snac_handle *h, int irk, double dt) { SNAC_TRY(h, do_rk_mom(h->ctx, irk, dt)) }
int snac_rk_scal(snac_handle *h, int irk, double dt, int iscal) { SNAC_TRY(h, do_rk_scal(h->ctx, irk, dt, iscal)) }
int snac_bounduvw(snac_handle *h, int is_correc) { SNAC_TRY(h, do_bounduvw(h->ctx, is_correc != 0)) }
int snac_boundp(snac_handle *h, int field) { SNAC_TRY(h, do_boundp(h->ctx, field)) }
int snac_set_inflow(snac_handle *h, int rank, int idir, int ib, const double *veln, const double *velt1, const double *velt2) {
  SNAC_TRY(h, set_inflow(h->ctx, rank, idir, ib, veln, velt1, velt2))
}
int snac_inflow(snac_handle *h, int is_correc) { SNAC_TRY(h, do_inflow(h->ctx, is_correc != 0)) }
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
  if (h) h->ctx.vel_epoch++;
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
      // MPI_ALLREDUCE(MIN) over every rank (src/chkdt.f90:52-53): the grids of all boxes of the run were recorded by
      // snac_add_box, so every process takes the minimum over all of them and arrives at the same dt
      double dlmin = 1e300;
      for (int r = 0; r < ctx.prm.nrank; ++r)
        for (int d = 0; d < 3; ++d) {
          const std::vector<double> &df = ctx.world_set[r] ? ctx.world_df[d][r] : std::vector<double>();
          for (double v : df) dlmin = std::min(dlmin, v);
        }
      double alpha_max = 0.0;
      for (int s = 0; s < ctx.prm.nscal; ++s) alpha_max = std::max(alpha_max, 1.0 / ctx.prm.alphai[s]);
      const double visc = std::max(ctx.prm.visc, alpha_max);
      dtmax = std::min(dtmax, 1.65 / 12.0 / visc * dlmin * dlmin);
    }
    *dt_cfl = dtmax;
  })
}
// out1d / out2d (src/output.f90:34-225): contribution of box `rank`, see output.cu
static void out_reduce(Ctx &ctx, int rank, int field, int which, int idir, const double *wa, const double *wb, double scale, double *out) {
  Box &b = ctx.box_of_rank(rank);
  bool alias = false;
  Field3 *f = field_of(ctx, b, field, alias);
  if (!f || !f->p) throw Error("out1d/out2d: unknown field");
  if (!wa || (which == 1 && !wb) || !out) throw Error("out1d/out2d: null argument");
  launch_out_reduce(ctx, b.g, f->p, which, idir, wa, wb, scale, out);
}
int snac_out1d(snac_handle *h, int rank, int field, int idir, const double *wa, const double *wb, double scale, double *p1d) {
  SNAC_TRY(h, out_reduce(h->ctx, rank, field, 1, idir, wa, wb, scale, p1d))
}
int snac_out2d(snac_handle *h, int rank, int field, int idir, const double *w, double scale, double *p2d) {
  SNAC_TRY(h, out_reduce(h->ctx, rank, field, 2, idir, w, nullptr, scale, p2d))
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
__global__ void empty_kernel() {}

int snac_profile_enable(snac_handle *h, int on) {
  SNAC_TRY(h, {
    SNAC_CUDA(cudaStreamSynchronize(h->ctx.stream));
    h->ctx.prof_on = on != 0;
    if (on) {
      for (int c = 0; c < PC_COUNT; ++c) { h->ctx.prof_ms[c] = 0.0; h->ctx.prof_n[c] = 0; h->ctx.prof_units[c] = 0.0; }
      // what an event pair measures around an empty launch on the idle stream (launch latency): every
      // profiled scope starts on an idle stream too, so this much of its time is not kernel time
      Ctx &c = h->ctx;
      double best = 1e30;
      for (int r = 0; r < 20; ++r) {
        cudaEventRecord(c.prof_ev[0], c.stream);
        empty_kernel<<<1, 32, 0, c.stream>>>();
        cudaEventRecord(c.prof_ev[1], c.stream);
        cudaEventSynchronize(c.prof_ev[1]);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c.prof_ev[0], c.prof_ev[1]);
        if (r >= 5 && ms < best) best = ms;
      }
      c.prof_overhead_ms = best < 1e29 ? best : 0.0;
    }
  })
}
int snac_profile_get(snac_handle *h, int cls, double *ms, long long *launches) {
  SNAC_TRY(h, {
    if (cls < 0 || cls >= PC_COUNT) throw Error("snac_profile_get: class out of range");
    if (ms) *ms = h->ctx.prof_ms[cls];
    if (launches) *launches = h->ctx.prof_n[cls];
  })
}
int snac_profile_units(snac_handle *h, int cls, double *points) {
  SNAC_TRY(h, {
    if (cls < 0 || cls >= PC_COUNT) throw Error("snac_profile_units: class out of range");
    if (points) *points = h->ctx.prof_units[cls];
  })
}
const char *snac_profile_name(int cls) {
  static const char *names[PC_COUNT] = {"rk_mom", "halo_bc", "fillps", "fft_fwd", "mg_smooth_l0", "mg_down_l0",
                                        "mg_up_l0", "mg_coarse_levels", "fft_bwd", "correc_updt_pressure", "chkdt_chkdiv",
                                        "direct_transform", "direct_tridiag"};
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
int snac_plane_solver(const snac_handle *h) { return h ? (h->ctx.direct_on ? 2 : 1) : -1; }
long long snac_launch_count(const snac_handle *h) { return h ? h->ctx.launches : -1; }
int snac_sync(snac_handle *h) {
  SNAC_TRY(h, {
    SNAC_CUDA(cudaStreamSynchronize(h->ctx.stream));
    check_peers(h->ctx);
  })
}
void *snac_stream(snac_handle *h) { return h ? (void *)h->ctx.stream : nullptr; }

}  // extern "C"
