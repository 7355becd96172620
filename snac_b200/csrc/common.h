// Internal declarations shared by the CUDA translation units of libsnacb200.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/snac_b200.h"

namespace snac {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define SNAC_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (expr);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      throw ::snac::Error(std::string(#expr) + " failed: " + cudaGetErrorString(e_) + " at " +   \
                          __FILE__ + ":" + std::to_string(__LINE__));                            \
  } while (0)

inline long round_up(long a, long b) { return (a + b - 1) / b * b; }

// Ghosted 3-D array in device memory, x fastest.  Element (i,j,k), i in [0,n1+1] (Fortran index
// lo-1+i), lives at p[i + sy*j + sz*k].  The allocation is padded so that the first interior
// element of every row, (1,j,k), is 128-byte aligned and sy is a multiple of 16 doubles.
struct Grid3 {
  int n[3] = {0, 0, 0};  // interior sizes
  long sy = 0, sz = 0;
  long alloc_elems() const { return 16 + sz * (n[2] + 2); }
  long st(int d) const { return d == 0 ? 1 : (d == 1 ? sy : sz); }
  static Grid3 make(int n1, int n2, int n3) {
    Grid3 g;
    g.n[0] = n1; g.n[1] = n2; g.n[2] = n3;
    g.sy = round_up(n1 + 2, 16);
    g.sz = g.sy * (n2 + 2);
    return g;
  }
};

// Owning device buffer of doubles (zero-initialised).
struct DevBuf {
  double *raw = nullptr;
  size_t elems = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : raw(o.raw), elems(o.elems) { o.raw = nullptr; o.elems = 0; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) { release(); raw = o.raw; elems = o.elems; o.raw = nullptr; o.elems = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    elems = n;
    if (n) {
      SNAC_CUDA(cudaMalloc(&raw, n * sizeof(double)));
      SNAC_CUDA(cudaMemset(raw, 0, n * sizeof(double)));
    }
  }
  void release() {
    if (raw) cudaFree(raw);
    raw = nullptr;
    elems = 0;
  }
};

// A field = buffer + pointer to its (0,0,0) ghost corner.
struct Field3 {
  DevBuf buf;
  double *p = nullptr;  // element (0,0,0)
  void alloc(const Grid3 &g) {
    buf.alloc(g.alloc_elems());
    p = buf.raw + 15;  // so that (1,0,0) is 128-byte aligned
  }
  void adopt(double *raw) { p = raw + 15; }   // memory owned by the communication arena
};

// 1-D device array uploaded from host.
struct Dev1 {
  DevBuf buf;
  const double *p() const { return buf.raw; }
  void upload(const std::vector<double> &h) {
    buf.alloc(h.size());
    if (!h.empty())
      SNAC_CUDA(cudaMemcpy(buf.raw, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
};

// ------------------------------------------------------------------------------------------
// halo copies and boundary-condition face fills (bound.cu)
// ------------------------------------------------------------------------------------------
struct CopyOp {  // dst plane <- src plane, both 2-D with extents (na, nb)
  const double *src;
  double *dst;
  long ssa, ssb, dsa, dsb;
  int na, nb;
};

enum FaceKind : int {
  FACE_CEN_LO = 0,   // q(0)   = factor + sgn*q(1)
  FACE_CEN_HI = 1,   // q(n+1) = factor + sgn*q(n)
  FACE_STAG_D_LO = 2,  // q(0) = factor
  FACE_STAG_D_HI = 3,  // q(n) = factor ; q(n+1) = q(n-1)
  FACE_STAG_N_LO = 4,  // q(0) = factor + q(1)
  FACE_STAG_N_HI = 5,  // q(n) = factor + q(n-1) ; q(n+1) = q(n)
  // inflow planes (src/bound.f90:438-477); `plane` holds the prescribed velocity per face cell
  FACE_INFLOW_N_LO = 6,  // q(0)   = plane
  FACE_INFLOW_N_HI = 7,  // q(n)   = plane
  FACE_INFLOW_T_LO = 8,  // q(0)   = -q(1) + 2 plane
  FACE_INFLOW_T_HI = 9   // q(n+1) = -q(n) + 2 plane
};

struct FaceOp {
  double *p;      // (0,0,0) corner of the field
  long sd, sa, sb;  // strides: normal, and the two in-plane directions
  int n, na, nb;  // interior size along the normal; in-plane extents (full ghost extents)
  int kind;
  double factor, sgn;
  const double *plane = nullptr;   // inflow kinds: dense (na, nb) plane, first index fastest
};

template <class T>
struct DevOps {  // device-resident array of ops
  T *d = nullptr;
  int count = 0;
  void upload(const std::vector<T> &h) {
    if (d) cudaFree(d);
    d = nullptr;
    count = (int)h.size();
    if (count) {
      SNAC_CUDA(cudaMalloc(&d, count * sizeof(T)));
      SNAC_CUDA(cudaMemcpy(d, h.data(), count * sizeof(T), cudaMemcpyHostToDevice));
    }
  }
  ~DevOps() { if (d) cudaFree(d); }
  DevOps() = default;
  DevOps(const DevOps &) = delete;
  DevOps &operator=(const DevOps &) = delete;
};

struct Ctx;

// Where the transforms of one box write the right-hand side to / read the solution from: one base
// pointer per owning process (the wavenumber axis is split into nsplit chunks of `chunk` planes).
struct FFTRoute {
  long st[3] = {0, 0, 0};          // strides of the physical directions in the solver arrays
  double *x[16] = {nullptr}, *b[16] = {nullptr};
  int chunk = 1, nsplit = 1;
  long off = 0;                    // first cell of the box inside its block's solver array
  double *xs[16][4] = {{nullptr}};  // x[p] for each of the arrays the owner's iterate can live in (MGSolver::xslot)
  bool cyclic = false;             // wavenumber q is solved by process q mod nproc (else: consecutive chunks)
};

// One ghosted array per local box taking part in an exchange.
struct ExField {
  int box;       // local box index
  double *p;     // (0,0,0) corner
  Grid3 g;
};

// Halo exchange of a group of fields (all local boxes), one direction at a time
// (updthalo, reference src/bound.f90:587-650).
struct NbTable {   // neighbour box ids (global, -1: none) of one local box in the array's own directions
  int nb[3][2];
};
struct Exchange {
  DevOps<CopyOp> local[3];
  DevOps<CopyOp> unpack[3];        // after the neighbours' pushes have arrived: landing buffers -> ghost planes (x-faces)
  unsigned mask[3] = {0, 0, 0};   // peer processes this exchange pushes to / receives from, per direction
  int key0 = -1;                  // arena key of the first exchanged array (remote neighbours), -1: local only
  // nbs == nullptr: the boxes' physical neighbour tables
  // rank_grids: the layout of the exchanged arrays on EVERY box of the run (multigrid levels, shortened staggered
  // ranges); default: Grid3::make of the box.  stage_x: x-faces of remote neighbours go through landing buffers.
  void build(Ctx &ctx, const std::vector<ExField> &fields, const std::vector<NbTable> *nbs = nullptr,
             const std::vector<int> *keys = nullptr, bool ids_local = false, const std::vector<Grid3> *rank_grids = nullptr,
             bool stage_x = true);
  void run(Ctx &ctx, int dir) const;      // one direction
  void run_all(Ctx &ctx) const;           // x, then y, then z (edges/corners propagate)
};

struct FaceFill {
  DevOps<FaceOp> ops[3];
  void upload(const std::vector<FaceOp> (&h)[3]) { for (int d = 0; d < 3; ++d) ops[d].upload(h[d]); }
  void run(Ctx &ctx) const;
};

void launch_copy_ops(Ctx &ctx, const DevOps<CopyOp> &ops, bool remote = false);
void launch_face_ops(Ctx &ctx, const DevOps<FaceOp> &ops);

// ------------------------------------------------------------------------------------------
// geometry passed by value to the stencil kernels
// ------------------------------------------------------------------------------------------
struct BoxGeom {
  int n[3];
  long sy, sz;
  // ghosted 1-D metric arrays (index 0..n+1) and their reciprocals
  const double *dc[3], *df[3], *rdc[3], *rdf[3];
};

// ------------------------------------------------------------------------------------------
// transforms (fft.cu)
// ------------------------------------------------------------------------------------------
enum R2RKind : int {  // names follow the FFTW r2r kinds the reference selects (src/fft.f90:96-149)
  K_R2HC = 0, K_HC2R = 1, K_DHT = 2, K_REDFT00 = 3, K_REDFT01 = 4, K_REDFT10 = 5, K_REDFT11 = 6,
  K_RODFT00 = 7, K_RODFT01 = 8, K_RODFT10 = 9, K_RODFT11 = 10
};

struct R2RPlan {
  int n = 0;       // box size along the axis
  int nt = 0;      // transform length (n - q)
  int axis = 0;
  int kind_fwd = 0, kind_bwd = 0;
  double normfft = 1.0;
  Dev1 mat_fwd, mat_bwd;  // dense nt x nt matrices (row-major [k][j]) for the generic path
  Dev1 twiddle;           // cos/sin tables of the fast path
  bool fast = false;
};

void find_fft(const char bc[2], bool centered, int &kf, int &kb, double norm[2]);
void r2r_plan_init(R2RPlan &pl, int axis, int n, const char bc[2], bool centered);
std::vector<double> r2r_eigenvalues(int n, const char bc[2], bool centered);
// in -> out along plan.axis over the interior of a ghosted box (in == out allowed); out *= scale
// in/out address the first interior element; ist/ost are the strides of the physical directions
// x,y,z in each array (in == out with equal strides is allowed); out *= scale
// in_split/out_split > 1: that side's transform axis is split into chunks of `chunk` indices, one
// base pointer per chunk (the owning process' slab, possibly in a peer's memory); cyclic: index t of the axis
// lives in piece t mod nsplit at position t div nsplit instead
void r2r_execute(Ctx &ctx, const R2RPlan &pl, int dir, const int n3[3], const double *const *in, int in_split,
                 const long ist[3], double *const *out, int out_split, const long ost[3], int chunk, double scale,
                 bool cyclic = false);

// ------------------------------------------------------------------------------------------
// multigrid (mg.cu)
// ------------------------------------------------------------------------------------------
struct MGLevelBox {
  Grid3 g;
  Field3 x, b;                 // x is owned except on level 0 where it may alias a context field
  double *xp = nullptr, *bp = nullptr;
  Dev1 am[3], ap[3], ad[3], w[3];   // ghosted 1-D coefficient arrays (index 0..n+1)
  Dev1 sam[3], sap[3], se[3], srw[3];  // scaled by the cell width: am/w, ap/w, ad/w, 1/w
  Dev1 tab[3];                         // packed {-am/w, -ap/w, ad/w, w, 1/w, 0, 0, 0} per index; entry 0 = outside cell (mg_ring.cu)
  std::vector<double> h_w[3];       // host copy of cell widths (for coarsening)
};

// one level of a plane-local batched 2-D problem as seen by the prefetch-ring kernels (mg_ring.cu)
struct RLevel {
  int n1, n2;
  int sy; long sz;           // row / plane strides of this level's arrays
  const double *ctab, *rtab; // packed coefficient tables of directions 0 and 1 (MGLevelBox::tab)
  int c1, c2;                // directions coarsened towards the next level
  int per1, per2;            // periodic wrap onto the box itself
};
void launch_ring_down(cudaStream_t st, const RLevel &L, int nsweep, const double *xin, double *xout, const double *b, double *cb, long csy,
                      long csz, const double *lamq, double alpha, double *norms, int nq, int want_b, const int *active, int nact, double xscale);
void launch_ring_up(cudaStream_t st, const RLevel &L, int nsweep, const double *xin, double *xout, const double *b, const double *cx, long csy,
                    long csz, const double *lamq, double alpha, int nq, const int *active, int nact);

// every level from `ls` down in one shared-memory kernel, one CTA per plane (mg_coarse.cu)
constexpr int CMAXLEV = 16;
struct CoarseLevelArg {
  int n1, n2, c1, c2, per1, per2;
  const double *t1, *t2;     // packed coefficient tables of directions 0 and 1 (MGLevelBox::tab)
};
struct CoarseArgs {
  CoarseLevelArg lev[CMAXLEV];
  int nlev;
  const double *b;           // right-hand side of the top level (ghosted layout, strides sy/sz)
  const double *xin;         // incoming iterate of the top level, or nullptr (zero)
  double *xout;
  long sy, sz;
  const double *lamq;
  double alpha;
  double *norms;             // only when the top level is level 0
  double xscale;             // factor applied to xin
  int nq, want_b, npre, npost, nbot;
  int npre_low, npost_low;   // sweeps per leg on the levels below the kernel's top level
  int adaptive;              // per-wavenumber depth of the hierarchy (mg_coarse.cu)
  const int *active;
};
size_t mg_coarse_smem(const CoarseArgs &a, bool single);
void launch_mg_coarse(cudaStream_t st, const CoarseArgs &a, bool single);

struct MGKernelBox {  // by-value kernel argument
  int n[3];
  long sy, sz;
  double *x;
  double *b;
  const double *am[3], *ap[3], *ad[3], *w[3];
  const double *lam;  // per-index eigenvalue along the uncoupled axis (or nullptr)
  int coupled[3];
  int fft_dir;        // -1 or the uncoupled direction
};

struct MGLevel {
  std::vector<MGLevelBox> boxes;   // one per local box
  Exchange ex;                     // halo exchange of x on this level
  int coarsen[3] = {0, 0, 0};      // directions coarsened going to the next level
};

// One box as seen by a linear solver.  With FFT_AXIS on, the solver works in a canonical layout:
// direction 2 is the transform (batch) direction and directions 0,1 are the two coupled ones in
// physical order; perm[sdir] gives the physical direction.
struct SBox {
  int n[3];
  int nb[3][2];                     // neighbour box id (global) or -1
  bool phys[3][2];                  // physical (non-periodic) boundary face
  double sgn[3][2];                 // boundary fold sign (src/solver.f90:71-94)
  std::vector<double> d1[3], d2[3]; // ghosted spacings: between unknowns / cell widths
  std::vector<double> lam;          // ghosted eigenvalues/dl^2 along direction 2 (batched only)
  double *x = nullptr, *b = nullptr; // level-0 arrays if they alias context fields (3-D solve)
  double *xalt[3] = {nullptr, nullptr, nullptr};   // several processes: arena arrays for MGSolver::xslot[..][1..3]
  long sy = 0, sz = 0;              // their strides when they differ from Grid3::make(n) (shortened staggered ranges)
  int qqq[3] = {0, 0, 0};           // 1: unknowns staggered along that direction (src/solver.f90:66-67)
};

// direct solver of separable planes (direct.cu)
struct TriArgs;
struct MGSolver;
struct DirectBox {
  R2RPlan plan;            // transform along the uniform direction of the plane
  Dev1 mu, stab;           // its eigenvalues / h^2; the 1-D operator of the other direction
  DevBuf ib, ie, ig;       // factorisation: 1/pivot (slab layout), coupling to the last unknown (cyclic), 1/g
  bool cyclic = false;
};
struct DirectSolver {
  bool on = false;
  int tdir = 0;            // plane direction that is transformed (0 or 1)
  std::vector<DirectBox> box;
  void setup(Ctx &ctx, MGSolver &S, int tdir);
  void solve(Ctx &ctx, MGSolver &S);
  TriArgs args(const MGSolver &S, int lb, double alpha) const;
};
int direct_pick(const std::vector<SBox> &sboxes, bool ids_local, const std::vector<int> &local_of_rank);

struct MGCoopState;   // mg_coop.cu: the general path as one cooperative launch per solve
MGCoopState *mg_coop_setup(Ctx &ctx, MGSolver &S, const bool coupled[3]);
void mg_coop_solve(Ctx &ctx, MGSolver &S, MGCoopState *st, int npre, int npost, int nbot);
void mg_coop_collect(Ctx &ctx, MGSolver &S, MGCoopState *st);

struct MGSolver {
  DirectSolver direct;
  MGCoopState *coop = nullptr;
  bool coop_tried = false;
  int fft_dir = -1;                // uncoupled direction in solver order (2 when batched) or -1
  int perm[3] = {0, 1, 2};         // solver direction -> physical direction
  std::vector<Field3> x0, b0;      // owned level-0 arrays (batched layout)
  std::vector<SBox> sboxes;
  bool centered[3] = {true, true, true};
  std::vector<std::unique_ptr<MGLevel>> levels;
  std::vector<Dev1> lam;           // per local box (ghosted along the fft axis), (lambda)/dl^2
  int nq = 1;                      // number of independent planes (wavenumbers)
  DevBuf norms;                    // [2*nq]: sum r^2, sum b^2 per plane
  std::vector<double> h_norms;
  bool ids_local = false;          // neighbour ids of the solver boxes are solver-box indices (batched: blocks)
  // several processes, boxes coupled across them (3-D solves, Helmholtz solves of IMPDIFF): the level arrays live in
  // the arena, the halo tables are built once the peers are mapped (connect), residual norms are summed over the
  // processes (every process takes the same decisions: the handshake sequences stay aligned)
  bool distributed = false;
  int key0 = -1;                   // arena key of the level-0 iterate
  std::vector<std::vector<Grid3>> rank_grids;   // [level][rank]: layout of that level's arrays on every box of the run
  void connect(Ctx &ctx);
  bool plane_local = false;        // fused shared-memory path (planes not split over boxes)
  int first_smem_level = 0;        // first level handled by the one-CTA-per-plane kernel
  bool coarse_single = false;      // that kernel computes in FP32 (only ever a coarse-grid correction, never level 0)
  std::vector<Field3> x1;          // second x buffer of the tiled levels [level*nbox + box]
  int *d_active = nullptr, *d_where = nullptr, *d_iters = nullptr, *d_count = nullptr;  // per-plane stopping flags
  int *h_count = nullptr;          // pinned: [0] the polled counter, then the histories below
  // planes still iterating after every cycle of the last solve, per caller-chosen key (the sub-step):
  // consecutive time steps behave alike, so the record of the previous solve sizes the launches of
  // this one cycle by cycle and tells how many cycles to enqueue before the first poll
  static constexpr int HIST_KEYS = 4, HIST_LEN = 64;
  int hist_key = 0;
  int hist_cur[HIST_KEYS] = {0, 0, 0, 0};   // which of the two records of a key is being written
  int hist_n[HIST_KEYS] = {0, 0, 0, 0};     // cycles in the complete record (0: none yet)
  int *hist(int key, int which) const { return h_count + 16 + ((key * 2 + which) * HIST_LEN); }
  double last_mean_iters = 0.0;
  int est_active = 0;              // planes still iterating at the last poll (sizes the work items of the next launches)
  void stats(Ctx &ctx, int *iters, double *relres, double *mean_iters);
  int last_iters = 0;
  double last_relres = 0.0;
  double alpha = 0.0;              // extra diagonal shift (IMPDIFF: alphai)
  int default_sweeps = 2;          // smoothing sweeps per leg when snac_params does not say: V(2,2) converges at ~0.13 per
                                   // cycle on isotropic grids (V(1,1): 0.27) and stays convergent on stretched / anisotropic ones
  double guess_scale = 1.0;        // factor applied to the initial guess (the previous solution) by the first cycle
  void setup(Ctx &ctx, std::vector<SBox> &&boxes, bool batched, int max_levels = 24, int key0 = -1);
  void solve(Ctx &ctx);
  const Grid3 &grid0(int lb) const { return levels[0]->boxes[lb].g; }
  // Plane-local path: four arrays can hold the level-0 iterate of a box -- the current one, the ping-pong
  // partner of the fused kernels, and one per key that is not current.  Initial guess per key: the
  // solution of the last solve with the same key (the same RK sub-step one time step earlier) is a
  // closer guess than the previous sub-step's, so each key keeps its own array and a solve starts by
  // making it the current one.  Every converged plane ends in the ping-pong partner, so the two trade
  // places at the end of a solve instead of a copy.  Both are exchanges of slot numbers; with several
  // processes the arrays live in the arena and every process applies the same exchanges to its
  // peers' addresses (the sequence of solves and keys is the same everywhere).
  static constexpr int NSLOT = 4;
  std::vector<std::array<double *, NSLOT>> xslot;   // per box
  std::vector<Field3> slot_own[NSLOT];               // arrays owned here (one process; allocated on first use)
  int slot_primary = 0, slot_scratch = 1;
  int slot_key[HIST_KEYS] = {-1, 2, 3, -1};          // where the array of a key rests (-1: it is the current one / none)
  int cur_key = 0;
  bool swapped = false;            // the level-0 iterate moved to the other buffer during the last solve (plane-local path)
  double *xcur(int lb) const { return levels[0]->boxes[lb].xp; }
  double *bcur(int lb) const { return levels[0]->boxes[lb].bp; }
};

// ------------------------------------------------------------------------------------------
// per-box state
// ------------------------------------------------------------------------------------------
struct Box {
  snac_box info;
  Grid3 g;
  std::vector<double> h_dc[3], h_df[3];
  Dev1 dc[3], df[3], rdc[3], rdf[3];
  Field3 u[2], v[2], w[2];  // double-buffered velocity (rk_mom writes the other parity)
  Field3 p, pp, rhs, po;
  Field3 stage;             // staging array for transfers of spectral fields (batched solves)
  Field3 du, dv, dw;        // d?dtrko
  Field3 s[SNAC_MAX_SCAL], ds[SNAC_MAX_SCAL];
  Field3 scr0, scr1;        // scratch of rk_scal (new RHS, diffusion term)
  double cbu[3][3][2];      // boundary RHS coefficients of the velocity Helmholtz problems [comp][idir][ib]
  double cbs[SNAC_MAX_SCAL][3][2];
  // Poisson RHS: weights (cell width, or 1 along the transform axis), 1/width along the transform
  // axis (else 1), and the boundary coefficient of init_bc_rhs per face
  Dev1 wq[3], gq[3];
  double cbp[3][2];
  Dev1 inflow[3][2][3];     // inflow planes [idir][ib][component u,v,w] (u_in/v_in/w_in%{x,y,z}, src/main.f90:232-240)
  bool is_inflow[3][2] = {{false, false}, {false, false}, {false, false}};
  // TMA tensor maps (CUtensorMap, 128 bytes each) of u,v,w per velocity buffer and of p ([0][3]) for the z-marching
  // rk_mom kernel (mom_tma.cu); tma_ok: they could be encoded
  alignas(64) unsigned char tmap[2][4][128];
  alignas(64) unsigned char tmap_hist[3][128];   // du, dv, dw (interior tiles)
  bool tma_ok = false;
  // ... and for the z-marching fillps / correc kernels (pois_tma.cu): per velocity buffer the tiles of u and v with
  // the column i-1 / the row j-1, the interior tiles of u,v,w ([0][3]: p), the tiles with halo of pp
  alignas(64) unsigned char tmap_fill[2][2][128];
  alignas(64) unsigned char tmap_int[2][4][128];
  alignas(64) unsigned char tmap_pp[128];
  bool tma_pois_ok = false;
  bool pp_is_po = false;    // 3-D solve: pp aliases the solver iterate
  double *pp_ptr() const { return pp_is_po ? po.p : pp.p; }
  BoxGeom geom() const;
};

// ------------------------------------------------------------------------------------------
// peer communication between the processes of one node (comm.cu): one arena per process holds every
// array a peer may write or read (field ghost layers, solver slabs, flags); arenas are mapped into
// each other with CUDA IPC so that halos and transposes are plain loads/stores over NVLink.
// ------------------------------------------------------------------------------------------
constexpr int MAXPROC = 16;
enum CommChannel : int { CH_READY = 0, CH_ARRIVED = 1, CH_FFT = 2, CH_SOLVED = 3, CH_REDUCE = 4, CH_COUNT = 5 };
enum ArenaKey : int {   // exported arrays: per box (rank) or per block
  AK_U0 = 0, AK_U1, AK_V0, AK_V1, AK_W0, AK_W1, AK_P, AK_PP, AK_S0 /* .. AK_S0+7 */, AK_SLAB_X = 24, AK_SLAB_B = 25,
  AK_SLAB_X1 = 26, AK_SLAB_K1 = 27, AK_SLAB_K2 = 28,  // the other arrays that can hold the solver's level-0 iterate (MGSolver::xslot)
  AK_PO = 16,          // iterate of the 3-D pressure solve (fft_axis = 0)
  AK_MGX = 32,         // + level: iterate of a coarse level of the 3-D pressure multigrid (levels 1 .. 23)
  AK_XSTAGE = 64   // + 2 * (key of the field) + side: dense landing buffer of an x-face pushed by the neighbour (bound.cu)
};
struct ArenaRec { int id, key; long long off; };   // id: box rank or block; off: bytes from the arena base
constexpr int MAXREC = 1024;
struct CommBlob {   // what every process publishes to the others (fixed size)
  int proc, nrec;
  cudaIpcMemHandle_t handle;
  ArenaRec rec[MAXREC];
};

struct PeerComm {
  int nproc = 1, me = 0;
  bool connected = false;
  char *arena = nullptr;
  size_t cap = 0, used = 0;
  char *peer[MAXPROC] = {nullptr};           // base of every process' arena in this address space
  std::vector<ArenaRec> recs[MAXPROC];       // published tables
  long long seq[CH_COUNT] = {0};
  // layout of the head of every arena: flags[CH_COUNT][MAXPROC] (long long), then red[MAXPROC][4] (double)
  static constexpr size_t FLAGS_BYTES = CH_COUNT * MAXPROC * sizeof(long long);
  static constexpr size_t HEAD_BYTES = 4096;
  // after the flags and the two reduction tables: [0] abort request written by a failing peer, [1] error raised by
  // this process' handshake kernel (comm.cu)
  static constexpr size_t ABORT_OFFSET = FLAGS_BYTES + 2 * MAXPROC * 4 * sizeof(double);
  static_assert(ABORT_OFFSET + 16 <= HEAD_BYTES, "arena head layout");
  void signal_abort();
  int error_state() const;
  void init(int nproc_, int me_, size_t bytes);
  double *alloc(size_t elems, int id, int key);   // zero-initialised, 256-byte aligned, recorded for export
  void export_blob(CommBlob &b) const;
  void connect(const CommBlob *blobs, int n);
  double *remote(int proc, int id, int key) const;  // address of a published array of process proc
  // all processes in `mask` signal each other on `ch` and wait (stream-ordered)
  void handshake(Ctx &ctx, int ch, unsigned mask);
  void release();
};

// per-kernel-class device timing (enabled by snac_profile_enable; serialises the stream)
enum ProfClass : int {
  PC_RK_MOM = 0, PC_HALO_BC, PC_FILLPS, PC_FFT_FWD, PC_MG_SMOOTH0, PC_MG_DOWN0, PC_MG_UP0, PC_MG_COARSE,
  PC_FFT_BWD, PC_CORREC, PC_CHK, PC_DIRECT_TRANSFORM, PC_DIRECT_TRIDIAG, PC_COUNT
};

struct Ctx {
  snac_params prm;
  bool prof_on = false;
  double prof_overhead_ms = 0.0;       // event-pair time around an empty launch on an idle stream (subtracted per scope)
  double prof_ms[PC_COUNT] = {0};
  long long prof_n[PC_COUNT] = {0};
  double prof_units[PC_COUNT] = {0};   // grid points processed (active planes only for the solver kernels)
  cudaEvent_t prof_ev[2] = {nullptr, nullptr};
  cudaEvent_t timer_ev[2] = {nullptr, nullptr};
  std::vector<int> owner;
  PeerComm comm;
  std::vector<snac_box> world;              // every box of the run (info only), indexed by rank
  std::vector<std::vector<double>> world_dc[3], world_df[3];   // their ghosted 1-D grids
  std::vector<char> world_set;
  std::vector<std::unique_ptr<Box>> boxes;  // local boxes
  std::vector<int> local_of_rank;           // rank -> local index or -1
  cudaStream_t stream = nullptr;
  bool committed = false;
  bool ready = false;                       // exchanges and routing built (after snac_comm_connect when nproc > 1)
  int nq = 1, q0 = 0, q1 = 1;               // wavenumbers of the run and the slab solved by this process (q1 - q0 of them)
  bool q_cyclic = false;                    // ... which are q0 + k nproc (cyclic) instead of q0 + k
  std::vector<int> block_index;             // block id -> solver box
  std::vector<std::array<int, 3>> blocks_lo;
  std::vector<int> blocks_id;               // solver box -> id under which its slabs are published (first member block)
  std::vector<FFTRoute> route;              // per local box
  // chkdt and chkdiv are one pass over u,v,w (K7): the reductions of the last pass are kept until the velocity changes
  // (the time loop calls chkdt and chkdiv back to back, src/main.f90:1048-1069)
  long long vel_epoch = 0, chk_epoch = -1;
  double chk_out[3] = {0, 0, 0};
  int par = 0;                              // which velocity buffer is current
  double prev_dtrk = 0.0;                   // sub-step size of the previous pressure solve (initial-guess scaling)
  long long launches = 0;
  std::string err;
  // exchanges / BC fills
  Exchange ex_uvw[2], ex_p, ex_pp;
  Exchange ex_uvwp[2];     // u,v,w,p together (end of a sub-step)
  FaceFill bc_uvw[2][2];   // [parity][is_correc]
  FaceFill bc_inflow[2][2];
  bool has_inflow = false, inflow_dirty = false;
  FaceFill bc_p, bc_pp;
  Exchange ex_s[SNAC_MAX_SCAL];
  FaceFill bc_s[SNAC_MAX_SCAL];
  // transforms and solvers
  R2RPlan plan_p;
  MGSolver psolver;
  bool direct_on = false;                           // the direct plane solver handles the pressure problem
  std::unique_ptr<MGSolver> usolver[3][2];          // IMPDIFF: [component][velocity buffer]
  std::unique_ptr<MGSolver> ssolver[SNAC_MAX_SCAL];
  double fs[SNAC_MAX_SCAL] = {0};                   // accumulated scalar forcing (src/main.f90:913)
  // reductions
  // snac_download_async: the field is packed into a staging array of its own in stream order, copied to the host on a
  // copy stream; a slot is free again once its copy has completed
  struct AsyncSlot { DevBuf stage; cudaEvent_t packed = nullptr, done = nullptr; };
  static constexpr int MAX_ASYNC = 32;
  AsyncSlot async_slot[MAX_ASYNC];
  cudaStream_t copy_stream = nullptr;
  DevBuf xfer;              // dense staging buffer of snac_upload / snac_download
  DevBuf red;               // scratch for chkdt/chkdiv
  double *h_red = nullptr;  // pinned
  double **d_slots = nullptr;
  Box &box_of_rank(int rank);
  // sum / max over the processes of up to 3 device values each (peer tables + one handshake; api.cu)
  void peer_reduce(const double *d_vals, int n, double *sum, double *mx);
};

struct ProfScope {
  Ctx &c;
  int cls;
  long long l0;
  ProfScope(Ctx &ctx, int cls_, double units = 0.0) : c(ctx), cls(cls_), l0(ctx.launches) {
    if (c.prof_on) { c.prof_units[cls] += units; cudaEventRecord(c.prof_ev[0], c.stream); }
  }
  ~ProfScope() {
    if (!c.prof_on) return;
    cudaEventRecord(c.prof_ev[1], c.stream);
    cudaEventSynchronize(c.prof_ev[1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c.prof_ev[0], c.prof_ev[1]);
    if (c.launches > l0) c.prof_ms[cls] += ms > c.prof_overhead_ms ? ms - c.prof_overhead_ms : 0.0;
    c.prof_n[cls] += c.launches - l0;
  }
};

// mom.cu / pois.cu entry points (one call handles one box)
void launch_rk_mom(Ctx &ctx, Box &b, int irk, double dt);
bool launch_rk_mom_tma(Ctx &ctx, Box &b, int irk, double dt);   // mom_tma.cu; false: not applicable, use launch_rk_mom
bool mom_tma_make_map(void *out128, const double *corner, const Grid3 &g, bool interior = false);
bool pois_tma_setup(Box &b);                                     // pois_tma.cu
bool launch_fillps_tma(Ctx &ctx, Box &b, double dtrk, bool do_bc);   // false: not applicable, use fillps_kernel
bool launch_correc_tma(Ctx &ctx, Box &b, double dtrk, double alpha, bool do_correc, bool do_updt);
void launch_fillps(Ctx &ctx, Box &b, double dtrk);
void launch_fillps_bc(Ctx &ctx, Box &b, double dtrk);
void launch_updt_rhs(Ctx &ctx, Box &b);
void launch_correc_updt(Ctx &ctx, Box &b, double dtrk, double alpha, bool do_correc, bool do_updt);
void launch_out_reduce(Ctx &ctx, const Grid3 &g, const double *field, int which, int idir, const double *wa, const double *wb,
                       double scale, double *host);   // output.cu: out1d / out2d plane averages of one box
void launch_chk(Ctx &ctx, Box &b, double *d_out /* [4]: dti_max, divtot, divmax, - */);
void launch_helm_rhs(Ctx &ctx, Box &b, const double *fld, const double *const wq[3], const double cb[3][2], const int nhi[3],
                     double alphai);
void launch_scal_rhs(Ctx &ctx, Box &b, int is, double alpha);
void launch_scal_update(Ctx &ctx, Box &b, int is, double f1, double f2, double f12, double *d_sum);
void launch_scal_finish(Ctx &ctx, Box &b, int is, double f12, double f);
void compute_bc_rhs(Ctx &ctx, Box &b);

}  // namespace snac
