// K4, general multi-box multigrid as ONE cooperative launch per solve.
//
// The general path of mg.cu (planes split over blocks that do not tile a rectangle, 3-D solves with FFT_AXIS = 0, the
// Helmholtz solves of IMPDIFF: reference solve_helmholtz / solve_n_helmholtz_2d, src/solver.f90:524-573,843-922) ran one
// launch per box, colour, level and direction and a host synchronisation per cycle: thousands of launches per solve on
// problems whose coarse levels hold a few hundred cells, i.e. bound by launch latency.  Here one persistent grid walks
// through the same V-cycle -- the same sweeps, restrictions, prolongations and block-interface copies in the same
// order, all boxes of a level in one phase -- with a grid-wide barrier where the launches used to end, evaluates the
// stopping rule (relative residual after pre-relaxation, per wavenumber, as PFMG does) on the device, and returns
// when it holds: one launch and no host round trip per solve.  A warp owns a row (box, j, k) of a level and its lanes
// stride along x, so the index arithmetic is paid per row and the accesses are coalesced.  The ghost planes across
// block interfaces and periodic wraps are refreshed by the copy tables of the level's Exchange (bound.cu), all three
// directions in one phase: the 7-point operators never read edge or corner ghosts, so the x -> y -> z order of the
// flow-field halos is not needed.  Used when every box a solver couples lives on this process; boxes coupled across
// processes keep the per-launch path (its handshakes are stream-ordered kernels).
#include <cooperative_groups.h>

#include <algorithm>
#include <cmath>

#include "common.h"
#include "mg_dev.cuh"

namespace cg = cooperative_groups;

namespace snac {

namespace {

#ifndef SNAC_MG_COOP_THREADS
#define SNAC_MG_COOP_THREADS 512
#endif
constexpr int COOP_THREADS = SNAC_MG_COOP_THREADS, COOP_WARPS = COOP_THREADS / 32;

struct CoopLevel {
  int box0, nbox;            // this level's entries of CoopArgs::boxes
  int co[3];                 // directions coarsened towards the next level
  const CopyOp *cp[3];       // block-interface / periodic ghost copies per direction
  int ncp[3];
  long rows;                 // interior rows (j,k) over all boxes
};

struct CoopBox {
  MGKernelBox k;
  long row0;                 // first row of this box among the rows of its level
  double *xraw;              // start of the iterate's allocation and its length (zeroed before a coarse-grid correction)
  long xelems;
};

struct CoopArgs {
  CoopLevel lev[24];
  int nlev, ltail;           // levels ltail .. nlev-1: few enough rows for one CTA (mg_coop_kernel)
  const CoopBox *boxes;
  double *norms;             // [nq] sum r^2, [nq] sum b^2
  int nq;
  double tol, alpha;
  int maxiter, npre, npost, nbot;
  int fft_dir;
  double *result;            // [0] cycles, [1] worst relative residual
};

struct Row { const CoopBox *b; int j, k; };

__device__ __forceinline__ Row row_of(const CoopArgs &a, const CoopLevel &L, long r) {
  int bi = L.box0;
  for (int q = 1; q < L.nbox; ++q)
    if (r >= a.boxes[L.box0 + q].row0) bi = L.box0 + q;
  const CoopBox *b = a.boxes + bi;
  const long lr = r - b->row0;
  const int n1 = b->k.n[1];
  return Row{b, (int)(lr % n1) + 1, (int)(lr / n1) + 1};
}

__device__ void coop_smooth(const CoopArgs &a, int l, int color, long gwarp, long nwarps, int lane) {
  const CoopLevel &L = a.lev[l];
  for (long r = gwarp; r < L.rows; r += nwarps) {
    const Row R = row_of(a, L, r);
    const MGKernelBox &B = R.b->k;
    const int par0 = (B.coupled[1] ? R.j : 0) + (B.coupled[2] ? R.k : 0);
    int i0 = 1, step = 32;
    if (B.coupled[0]) { i0 = ((color - par0 - 1) & 1) ? 2 : 1; step = 64; }
    else if ((par0 & 1) != color) continue;
    const double wy = __ldg(B.w[1] + R.j), wz = __ldg(B.w[2] + R.k);
    for (int i = i0 + (B.coupled[0] ? 2 * lane : lane); i <= B.n[0]; i += step) {
      const long c = i + B.sy * R.j + B.sz * R.k;
      const double wx = __ldg(B.w[0] + i);
      const double diag = mg_diag(B, i, R.j, R.k, wx, wy, wz, a.alpha);
      if (diag == 0.0) continue;
      B.x[c] = (B.b[c] - mg_offdiag(B, B.x, c, i, R.j, R.k, wx, wy, wz)) / diag;
    }
  }
}

__device__ void coop_exchange(const CoopArgs &a, int l, long gtid, long gsize) {
  const CoopLevel &L = a.lev[l];
  for (int d = 0; d < 3; ++d)
    for (int o = 0; o < L.ncp[d]; ++o) {
      const CopyOp op = L.cp[d][o];
      const long tot = (long)op.na * op.nb;
      for (long t = gtid; t < tot; t += gsize) {
        const int ia = (int)(t % op.na), ib = (int)(t / op.na);
        op.dst[ia * op.dsa + ib * op.dsb] = op.src[ia * op.ssa + ib * op.ssb];
      }
    }
}

// residual of level l summed over the children of each coarse cell -> right-hand side of level l+1, whose iterate is
// zeroed; optionally the per-plane sums of r^2 (and b^2).  On the last level only the norms are formed.
__device__ void coop_restrict(const CoopArgs &a, int l, bool norms, bool want_b, long gwarp, long nwarps, int lane, long gtid, long gsize) {
  const CoopLevel &F = a.lev[l];
  const bool has_coarse = l + 1 < a.nlev;
  const CoopLevel &C = a.lev[has_coarse ? l + 1 : l];
  if (has_coarse)
    for (int q = 0; q < C.nbox; ++q) {
      const CoopBox &cb = a.boxes[C.box0 + q];
      for (long t = gtid; t < cb.xelems; t += gsize) cb.xraw[t] = 0.0;
    }
  const int c0 = has_coarse ? F.co[0] : 0, c1 = has_coarse ? F.co[1] : 0, c2 = has_coarse ? F.co[2] : 0;
  for (long r = gwarp; r < C.rows; r += nwarps) {
    const Row R = row_of(a, C, r);
    const MGKernelBox &Fb = a.boxes[F.box0 + (int)(R.b - (a.boxes + C.box0))].k;
    const MGKernelBox &Cb = R.b->k;
    const int J = R.j, K = R.k;
    const int j0 = c1 ? 2 * J - 1 : J, j1 = c1 ? min(2 * J, Fb.n[1]) : J;
    const int k0 = c2 ? 2 * K - 1 : K, k1 = c2 ? min(2 * K, Fb.n[2]) : K;
    double r2 = 0.0, b2 = 0.0;
    for (int I = 1 + lane; I <= Cb.n[0]; I += 32) {
      const int i0 = c0 ? 2 * I - 1 : I, i1 = c0 ? min(2 * I, Fb.n[0]) : I;
      double rsum = 0.0;
      for (int k = k0; k <= k1; ++k)
        for (int j = j0; j <= j1; ++j)
          for (int i = i0; i <= i1; ++i) {
            const long c = i + Fb.sy * j + Fb.sz * k;
            const double wx = __ldg(Fb.w[0] + i), wy = __ldg(Fb.w[1] + j), wz = __ldg(Fb.w[2] + k);
            const double bb = Fb.b[c];
            const double res = bb - (mg_diag(Fb, i, j, k, wx, wy, wz, a.alpha) * Fb.x[c] + mg_offdiag(Fb, Fb.x, c, i, j, k, wx, wy, wz));
            rsum += res; r2 += res * res; b2 += bb * bb;
          }
      if (has_coarse) Cb.b[I + Cb.sy * J + Cb.sz * K] = rsum;
    }
    if (norms) {
      for (int o = 16; o > 0; o >>= 1) {
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        b2 += __shfl_xor_sync(0xffffffffu, b2, o);
      }
      if (lane == 0) {
        const int q = a.fft_dir < 0 ? 0 : K - 1;   // batched problems: the planes are indexed by direction 2, never coarsened
        atomicAdd(a.norms + q, r2);
        if (want_b) atomicAdd(a.norms + a.nq + q, b2);
      }
    }
  }
}

__device__ void coop_prolong(const CoopArgs &a, int l, long gwarp, long nwarps, int lane) {
  const CoopLevel &F = a.lev[l];
  const CoopLevel &C = a.lev[l + 1];
  for (long r = gwarp; r < F.rows; r += nwarps) {
    const Row R = row_of(a, F, r);
    const MGKernelBox &B = R.b->k;
    const MGKernelBox &Cb = a.boxes[C.box0 + (int)(R.b - (a.boxes + F.box0))].k;
    const int J = F.co[1] ? (R.j + 1) >> 1 : R.j, K = F.co[2] ? (R.k + 1) >> 1 : R.k;
    const double *cx = Cb.x + Cb.sy * J + Cb.sz * K;
    double *x = B.x + B.sy * R.j + B.sz * R.k;
    for (int i = 1 + lane; i <= B.n[0]; i += 32) x[i] += cx[F.co[0] ? (i + 1) >> 1 : i];
  }
}

// x = 0 on the planes whose right-hand side vanishes (HYPRE returns x = 0 for b = 0)
__device__ void coop_zero_planes(const CoopArgs &a, long gwarp, long nwarps, int lane) {
  const CoopLevel &L = a.lev[0];
  for (long r = gwarp; r < L.rows; r += nwarps) {
    const Row R = row_of(a, L, r);
    const MGKernelBox &B = R.b->k;
    const int q = a.fft_dir < 0 ? 0 : R.k - 1;
    if (a.norms[a.nq + q] != 0.0) continue;
    double *x = B.x + B.sy * R.j + B.sz * R.k;
    for (int i = 1 + lane; i <= B.n[0]; i += 32) x[i] = 0.0;
  }
}

// who works on a phase: the whole grid (phases end at a grid-wide barrier) or, on the coarse tail of the hierarchy,
// one CTA (phases end at __syncthreads: a level of a few hundred cells is done before a grid barrier would return)
struct Team { long tid, size, warp, nwarps; };

template <bool GRID>
struct Exec {
  const CoopArgs &a;
  cg::grid_group &grid;
  Team T;
  int lane;
  __device__ __forceinline__ void sync() { if (GRID) grid.sync(); else __syncthreads(); }
  __device__ void exchange(int l) {
    if (a.lev[l].ncp[0] + a.lev[l].ncp[1] + a.lev[l].ncp[2] == 0) return;   // (the same on every thread)
    coop_exchange(a, l, T.tid, T.size);
    sync();
  }
  __device__ void sweeps(int l, int n, bool reverse) {
    for (int s = 0; s < n; ++s)
      for (int c = 0; c < 2; ++c) {
        coop_smooth(a, l, reverse ? 1 - c : c, T.warp, T.nwarps, lane);
        sync();
        exchange(l);
      }
  }
  __device__ void restrict_from(int l, bool norms, bool want_b) {
    coop_restrict(a, l, norms, want_b, T.warp, T.nwarps, lane, T.tid, T.size);
    sync();
  }
  // one level of the downward leg (the bottom sweeps on the last level) / of the upward leg
  __device__ void down(int l) {
    if (l == a.nlev - 1) { sweeps(l, a.nbot, false); return; }
    sweeps(l, a.npre, false);
    restrict_from(l, false, false);
  }
  __device__ void up(int l) {
    coop_prolong(a, l, T.warp, T.nwarps, lane);
    sync();
    exchange(l);
    sweeps(l, a.npost, true);
  }
};

__global__ void __launch_bounds__(COOP_THREADS) mg_coop_kernel(const CoopArgs a) {
  cg::grid_group grid = cg::this_grid();
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gsize = (long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  __shared__ double s_w[COOP_WARPS];
  __shared__ int s_f[COOP_WARPS][2];
  Exec<true> G{a, grid, Team{gtid, gsize, gtid >> 5, gsize >> 5}, lane};
  Exec<false> B{a, grid, Team{threadIdx.x, blockDim.x, threadIdx.x >> 5, blockDim.x >> 5}, lane};
  const int lt = a.ltail;   // levels lt .. nlev-1 are walked by CTA 0 alone (nlev: none)

  G.exchange(0);
  int it = 0;
  double worst = 0.0;
  for (;; ++it) {
    G.sweeps(0, a.nlev == 1 ? a.nbot : a.npre, false);
    G.restrict_from(0, true, it == 0);
    // stopping rule (the same numbers, the same arithmetic on every CTA: every thread takes the same branch)
    double w = 0.0;
    int notconv = 0, anyzero = 0;
    for (int q = threadIdx.x; q < a.nq; q += blockDim.x) {
      const double r2 = a.norms[q], b2 = a.norms[a.nq + q];
      if (b2 == 0.0) { anyzero = 1; continue; }
      const double rel = sqrt(r2 / b2);
      w = fmax(w, rel);
      if (!(rel < a.tol)) notconv = 1;
    }
    for (int o = 16; o > 0; o >>= 1) {
      w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
      notconv |= __shfl_xor_sync(0xffffffffu, notconv, o);
      anyzero |= __shfl_xor_sync(0xffffffffu, anyzero, o);
    }
    if (lane == 0) { s_w[threadIdx.x >> 5] = w; s_f[threadIdx.x >> 5][0] = notconv; s_f[threadIdx.x >> 5][1] = anyzero; }
    __syncthreads();
    w = 0.0; notconv = 0; anyzero = 0;
    for (int q = 0; q < COOP_WARPS; ++q) { w = fmax(w, s_w[q]); notconv |= s_f[q][0]; anyzero |= s_f[q][1]; }
    __syncthreads();
    worst = w;
    const bool conv = !notconv;
    if (it == 0 && anyzero) {
      coop_zero_planes(a, G.T.warp, G.T.nwarps, lane);
      grid.sync();
      G.exchange(0);
      if (worst > 0.0) G.restrict_from(0, false, false);   // the coarse right-hand side used the old iterate there
    }
    if ((it > 0 && conv) || it >= a.maxiter || worst == 0.0 || (a.nlev == 1 && conv)) break;
    grid.sync();   // every CTA has read the norms: the sums of r^2 start from zero again
    if (blockIdx.x == 0)
      for (int q = threadIdx.x; q < a.nq; q += blockDim.x) a.norms[q] = 0.0;
    if (a.nlev > 1) {
      for (int l = 1; l < lt; ++l) G.down(l);
      if (lt < a.nlev) {
        if (blockIdx.x == 0) {
          for (int l = lt; l < a.nlev; ++l) B.down(l);
          for (int l = a.nlev - 2; l >= lt; --l) B.up(l);
        }
        grid.sync();
      }
      for (int l = min(lt, a.nlev - 1) - 1; l >= 0; --l) G.up(l);
    }
  }
  if (gtid == 0) { a.result[0] = (double)it; a.result[1] = worst; }
}

}  // namespace

struct MGCoopState {
  CoopArgs args;
  DevOps<CoopBox> boxes;
  DevBuf result;
  int grid = 0;
  double *h_result = nullptr;   // pinned copy of `result`
  bool pending = false;
};

void mg_coop_release(MGCoopState *s) {
  if (!s) return;
  if (s->h_result) cudaFreeHost(s->h_result);
  delete s;
}

// builds the device tables; nullptr when the path does not apply (boxes coupled across processes, a device without
// cooperative launches, more than 24 levels)
MGCoopState *mg_coop_setup(Ctx &ctx, MGSolver &S, const bool coupled[3]) {
  const bool off = getenv("SNAC_MG_COOP") && atoi(getenv("SNAC_MG_COOP")) == 0;   // A/B switch and test hook: one launch per box, colour and level
  if (off || S.distributed || S.levels.size() > 24 || S.fft_dir == 0 || S.fft_dir == 1) return nullptr;
  for (auto &L : S.levels)
    for (int d = 0; d < 3; ++d)
      if (L->ex.mask[d] || L->ex.unpack[d].count) return nullptr;
  int dev = 0, coop = 0, sms = 0, per_sm = 0;
  SNAC_CUDA(cudaGetDevice(&dev));
  SNAC_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  SNAC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SNAC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mg_coop_kernel, COOP_THREADS, 0));
  if (!coop || per_sm < 1) return nullptr;
  std::unique_ptr<MGCoopState> st(new MGCoopState);
  CoopArgs &a = st->args;
  std::vector<CoopBox> hb;
  a.nlev = (int)S.levels.size();
  long rows0 = 0;
  for (int l = 0; l < a.nlev; ++l) {
    MGLevel &L = *S.levels[l];
    CoopLevel &cl = a.lev[l];
    cl.box0 = (int)hb.size(); cl.nbox = (int)L.boxes.size();
    cl.rows = 0;
    for (int d = 0; d < 3; ++d) {
      cl.co[d] = L.coarsen[d];
      cl.cp[d] = coupled[d] ? L.ex.local[d].d : nullptr;
      cl.ncp[d] = coupled[d] ? L.ex.local[d].count : 0;
    }
    for (size_t lb = 0; lb < L.boxes.size(); ++lb) {
      MGLevelBox &B = L.boxes[lb];
      CoopBox cb;
      for (int d = 0; d < 3; ++d) {
        cb.k.n[d] = B.g.n[d];
        cb.k.am[d] = B.am[d].p(); cb.k.ap[d] = B.ap[d].p(); cb.k.ad[d] = B.ad[d].p(); cb.k.w[d] = B.w[d].p();
        cb.k.coupled[d] = coupled[d];
      }
      cb.k.sy = B.g.sy; cb.k.sz = B.g.sz;
      cb.k.x = B.xp; cb.k.b = B.bp;
      cb.k.fft_dir = S.fft_dir;
      cb.k.lam = S.fft_dir >= 0 ? S.lam[lb].p() : nullptr;
      cb.row0 = cl.rows;
      cl.rows += (long)B.g.n[1] * B.g.n[2];
      cb.xraw = l > 0 ? B.xp - 15 : nullptr;   // (level 0 is never zeroed: it may alias a flow field with its own strides)
      cb.xelems = l > 0 ? B.g.alloc_elems() : 0;
      hb.push_back(cb);
    }
    if (l == 0) rows0 = cl.rows;
  }
  // the coarse tail: from the first level (below level 0) with at most four rows per warp of a CTA
  a.ltail = a.nlev;
  for (int l = a.nlev - 1; l >= 1 && a.lev[l].rows <= 4 * COOP_WARPS; --l) a.ltail = l;
  st->boxes.upload(hb);
  a.boxes = st->boxes.d;
  a.norms = S.norms.raw;
  a.nq = S.nq;
  a.fft_dir = S.fft_dir;
  st->result.alloc(2);
  a.result = st->result.raw;
  // the grid is no wider than level 0 needs (a grid-wide barrier costs more the more CTAs take part); all CTAs must be
  // resident at once
  const long want = std::max(1L, (rows0 + 2 * COOP_WARPS - 1) / (2 * COOP_WARPS));   // two rows of level 0 per warp
  // one CTA per SM: a grid-wide barrier costs more the more CTAs take part, and the phases are short (128^3 cavity: 8.0 ms per
  // step with two CTAs per SM, 6.4 ms with one); SNAC_MG_COOP_GRID: tuning hook, the number of CTAs
  long cap = sms;
  if (getenv("SNAC_MG_COOP_GRID")) cap = std::min<long>((long)sms * per_sm, std::max(1, atoi(getenv("SNAC_MG_COOP_GRID"))));
  st->grid = (int)std::min<long>(want, cap);
  SNAC_CUDA(cudaMallocHost(&st->h_result, 2 * sizeof(double)));
  st->h_result[0] = st->h_result[1] = 0.0;
  return st.release();
}

void mg_coop_solve(Ctx &ctx, MGSolver &S, MGCoopState *st, int npre, int npost, int nbot) {
  CoopArgs &a = st->args;
  a.tol = ctx.prm.hypre_tol;
  a.maxiter = std::max(ctx.prm.hypre_maxiter, 1);
  a.alpha = S.alpha;
  a.npre = npre; a.npost = npost; a.nbot = nbot;
  for (int lb = 0; lb < a.lev[0].nbox; ++lb)   // (the level-0 arrays of a solver are fixed on this path; checked, not assumed)
    if (S.levels[0]->boxes[lb].xp == nullptr) throw Error("multigrid: level-0 iterate missing");
  SNAC_CUDA(cudaMemsetAsync(S.norms.raw, 0, 3 * (size_t)S.nq * sizeof(double), ctx.stream));
  void *params[1] = {&a};
  SNAC_CUDA(cudaLaunchCooperativeKernel((const void *)mg_coop_kernel, dim3(st->grid), dim3(COOP_THREADS), params, 0, ctx.stream));
  ctx.launches++;
  SNAC_CUDA(cudaMemcpyAsync(st->h_result, a.result, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  st->pending = true;
}

// the cycle count and the residual of the last solve, read back only when somebody asks (snac_solver_stats)
void mg_coop_collect(Ctx &ctx, MGSolver &S, MGCoopState *st) {
  if (!st || !st->pending) return;
  SNAC_CUDA(cudaStreamSynchronize(ctx.stream));
  S.last_iters = (int)st->h_result[0];
  S.last_mean_iters = st->h_result[0];
  S.last_relres = st->h_result[1];
  st->pending = false;
}

}  // namespace snac
