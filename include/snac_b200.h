/* snac_b200.h -- C ABI of the B200-native pressure-correction pipeline of SNaC.
 *
 * Drop-in boundary: the body of the RK3 sub-step loop of the reference driver
 * (reference src/main.f90:906-1033) plus its one-time set-up calls (src/main.f90:726-886).
 * The reference binds its native dependencies with ISO_C_BINDING `bind(C)` interfaces
 * (src/fftw.f90:9-28); this header is what a `bind(C)` interface module for the GPU path would
 * mirror.  Everything is `extern "C"`, plain pointers and sizes.
 *
 * Conventions
 *   - "box" = what the reference calls an MPI rank: one Cartesian sub-box `lo:hi` of one block
 *     (src/initmpi.f90:117-132).  A process (one per GPU) may own several boxes.
 *   - small index arrays are C-ordered: [idir][ib] with idir = 0,1,2 (x,y,z) and ib = 0 (lower
 *     face), 1 (upper face); the Fortran arrays are (0:1,1:3), i.e. the transpose.
 *   - host field arrays are laid out exactly like the reference's allocations
 *     `u(lo(1)-1:hi(1)+1, lo(2)-1:hi(2)+1, lo(3)-1:hi(3)+1)` (src/main.f90:226-231): column-major,
 *     one ghost layer, the pointer addresses element (lo-1,lo-1,lo-1).
 *   - every entry point returns 0 on success and a negative value on error;
 *     `snac_last_error` returns the message.  There is no CPU fallback: without a CUDA device
 *     `snac_create` fails.
 */
#ifndef SNAC_B200_H
#define SNAC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SNAC_ABI_VERSION 2

#if defined(__GNUC__)
#define SNAC_API __attribute__((visibility("default")))
#else
#define SNAC_API
#endif

typedef struct snac_handle snac_handle;

/* field identifiers for upload/download/boundp */
enum {
  SNAC_FIELD_U = 0,   /* u(lo-1:hi+1,...)  src/main.f90:226 */
  SNAC_FIELD_V = 1,
  SNAC_FIELD_W = 2,
  SNAC_FIELD_P = 3,   /* pressure p */
  SNAC_FIELD_PP = 4,  /* pressure correction pp in physical space (after the backward transform) */
  SNAC_FIELD_RHS = 5, /* Poisson right-hand side: the reference's pp between fillps and the solve */
  SNAC_FIELD_PO = 6,  /* solver iterate / initial guess po (src/main.f90:1010,1022); spectral space when fft_axis > 0 */
  SNAC_FIELD_DUDTRKO = 7,
  SNAC_FIELD_DVDTRKO = 8,
  SNAC_FIELD_DWDTRKO = 9,
  SNAC_FIELD_RHS_HAT = 10, /* transformed right-hand side (index along the transform axis = wavenumber) */
  SNAC_FIELD_S0 = 16,      /* scalar i  = SNAC_FIELD_S0 + i        */
  SNAC_FIELD_DSDTRKO0 = 32 /* its RK history = SNAC_FIELD_DSDTRKO0 + i */
};

#define SNAC_MAX_SCAL 8

/* Global (per run) parameters: what the hot path reads from mod_param (src/param.f90:23-45,
 * 88-106) plus the reference's compile-time switches (configs/defaults/build-default.conf:12-19)
 * turned into run-time values. */
typedef struct snac_params {
  int abi_version;        /* SNAC_ABI_VERSION */
  int device;             /* CUDA device ordinal of this process */
  int nrank;              /* number of boxes in the run (reference: MPI_COMM_WORLD size) */
  int nproc;              /* number of processes (GPUs) */
  int myproc;             /* this process */
  const int *rank_owner;  /* [nrank] process owning each box (NULL: all owned by process 0) */
  int periods[3];         /* global extent in periodic directions, else 0 (src/initmpi.f90:107-113) */
  int fft_axis;           /* 0: 3-D solve; 1,2,3: transforms along x,y,z (FFT_AXIS build option) */
  int impdiff;            /* IMPDIFF build option; with fft_axis > 0 the Helmholtz systems of solve_impdiff_field_fft
                           * (src/solve_helmholtz.f90:32-100) are solved in 3-D: the same linear systems */
  int boussinesq;         /* BOUSSINESQ_BUOYANCY build option */
  int nscal;
  double visc;            /* 1/visci (src/param.f90:175) */
  double bforce[3];
  double gacc[3];
  double beta;
  double vol_all;         /* src/param.f90:239-240 */
  double alphai[SNAC_MAX_SCAL];   /* scalar inverse diffusivities (&scalar) */
  double ssource[SNAC_MAX_SCAL];
  int is_sforced[SNAC_MAX_SCAL];
  double scalf[SNAC_MAX_SCAL];
  double hypre_tol;       /* relative residual tolerance of the linear solves (src/param.f90:15-17) */
  int hypre_maxiter;
  int hypre_solver_i;     /* accepted for compatibility; the solver is geometric multigrid or the direct plane solver */
  /* multigrid tuning (0 = default) */
  int mg_npre, mg_npost, mg_nbottom;
  int mg_generic;         /* 1: always use the general multi-box multigrid kernels */
  /* solver of the per-wavenumber 2-D problems (fft_axis > 0): 0 = auto (the direct solver where every plane is one
   * rectangle with a uniform periodic/Neumann/Dirichlet direction, multigrid otherwise), 1 = multigrid, 2 = direct
   * (snac_commit fails where the planes are not separable) */
  int plane_solver;
} snac_params;

/* Per-box description: the outputs of initmpi / initgrid / read_input that the hot path uses
 * (src/initmpi.f90:9-10, src/initgrid.f90:49-127, src/param.f90:46-57). */
typedef struct snac_box {
  int rank;               /* 0-based box id (reference myid) */
  int block;              /* 0-based block id */
  int lo[3], hi[3];       /* global indices, 1-based like the reference */
  int nb[3][2];           /* neighbour box id or -1 (MPI_PROC_NULL) */
  int is_bound[3][2];
  char cbcvel[3][3][2];   /* [ivel][idir][ib] : 'D','N','F' (block_cbcvel(ib,idir,ivel)) */
  double bcvel[3][3][2];
  char cbcpre[3][2];
  double bcpre[3][2];
  char cbcscal[SNAC_MAX_SCAL][3][2];
  double bcscal[SNAC_MAX_SCAL][3][2];
  /* 1-D metric arrays, each of length n+2 (index lo-1..hi+1), AFTER bound_grid */
  const double *dxc, *dxf, *dyc, *dyf, *dzc, *dzf;
  double dl[3][2];        /* d?c_g(lo_g-1), d?c_g(hi_g)  (src/main.f90:726-728) */
  double dlf[3][2];       /* d?f_g(lo_g),   d?f_g(hi_g)  (src/main.f90:771-773; IMPDIFF) */
} snac_box;

/* ---- life cycle ------------------------------------------------------------------------- */
SNAC_API int snac_create(const snac_params *params, snac_handle **out);
SNAC_API int snac_add_box(snac_handle *h, const snac_box *box);
/* Builds device state for all added boxes: fields, boundary/halo tables, transform plans
 * (fftini, src/fft.f90:10-69), boundary RHS (init_bc_rhs, src/solver.f90:44-157), multigrid
 * hierarchies (replacing init_matrix_3d, init_matrix_2d, create_solver and setup_solver,
 * src/solver.f90:158-484). */
SNAC_API int snac_commit(snac_handle *h);
/* Several processes (one per GPU of a node; params.nproc > 1): every process must add EVERY box of
 * the run (those it does not own are recorded as descriptions only), commit, publish a blob of
 * snac_comm_blob_size() bytes, gather the blobs of all processes in process order by its own means
 * (MPI_Allgather in the reference's setting, src/initmpi.f90:43-248 gathers the same information)
 * and connect: halos (updthalo, src/bound.f90:587-650) and the wavenumber transposes of the solver
 * (cf. src/solver.f90:1244-1371) then go over NVLink through CUDA-IPC mapped peer memory. */
#include <stddef.h>
SNAC_API size_t snac_comm_blob_size(void);
SNAC_API int snac_comm_export(snac_handle *h, void *blob);
SNAC_API int snac_comm_connect(snac_handle *h, const void *blobs, int nblobs);
SNAC_API int snac_finalize(snac_handle *h);
SNAC_API const char *snac_last_error(const snac_handle *h);   /* h may be NULL: error of snac_create */

/* ---- field transfer (host layout: see header comment) ------------------------------------ */
SNAC_API int snac_upload(snac_handle *h, int rank, int field, const double *host);
SNAC_API int snac_download(snac_handle *h, int rank, int field, double *host);
/* Outputs and checkpoints without stopping the time loop (the reference writes them with blocking MPI-IO from the host
 * arrays, src/load.f90:84-187, src/output.f90:227-257): snac_download_async packs what the field holds NOW into a device
 * staging array in stream order, starts the device -> host copy on a copy stream and returns a ticket; the caller goes on
 * stepping.  snac_download_wait blocks until that copy has landed; it may be called from another host thread (a file
 * writer).  `host`: layout of snac_download; pinned memory (snac_host_alloc) lets the copy overlap with the kernels. */
SNAC_API int snac_download_async(snac_handle *h, int rank, int field, double *host, int *ticket);
SNAC_API int snac_download_wait(snac_handle *h, int ticket);
SNAC_API void *snac_host_alloc(size_t bytes);
SNAC_API void snac_host_free(void *p);

/* ---- fine-grained entry points, 1:1 with the reference subroutines -------------------------
 * Each acts on all boxes owned by the process. */
SNAC_API int snac_rk_mom(snac_handle *h, int irk, double dt);            /* rk_mom, src/rk.f90:7-128; irk=1..3 */
SNAC_API int snac_rk_scal(snac_handle *h, int irk, double dt, int iscal);/* rk_scal + bulk_forcing_s, src/rk.f90:130-204 */
SNAC_API int snac_bounduvw(snac_handle *h, int is_correc);               /* bounduvw, src/bound.f90:8-66 */
SNAC_API int snac_boundp(snac_handle *h, int field);                     /* boundp, src/bound.f90:68-106 */
/* inflow planes of one face of box `rank` (u_in/v_in/w_in%{x,y,z}(:,:,ib), src/main.f90:232-240,594-632):
 * three column-major planes with ghosts, (n_a+2) x (n_b+2), a < b the in-plane directions; veln is the
 * component normal to the face, velt1/velt2 the other two in x<y<z order (init_inflow, src/initflow.f90:369-429).
 * idir = 0..2, ib = 0 (lower) or 1 (upper); the face must be a physical boundary with cbcvel 'D' for the
 * normal component (is_bound_inflow, src/main.f90:596). */
SNAC_API int snac_set_inflow(snac_handle *h, int rank, int idir, int ib, const double *veln, const double *velt1, const double *velt2);
SNAC_API int snac_inflow(snac_handle *h, int is_correc);                 /* inflow, src/bound.f90:438-477; no-op without planes */
SNAC_API int snac_fillps(snac_handle *h, double dtrk);                   /* fillps, src/fillps.f90:7-54 -> RHS */
SNAC_API int snac_updt_rhs(snac_handle *h);                              /* updt_rhs, src/bound.f90:543-585 on RHS */
SNAC_API int snac_fft(snac_handle *h, int dir);  /* fft, src/fft.f90:85-94: dir=+1 RHS -> RHS_HAT; dir=-1 PO -> PP, times normfft (src/main.f90:1017-1020) */
SNAC_API int snac_solve_n_helmholtz_2d(snac_handle *h);                  /* src/solver.f90:843-922: RHS,PO -> PO */
SNAC_API int snac_solve_helmholtz(snac_handle *h);                       /* src/solver.f90:524-573: RHS,PO -> PO (=PP) */
SNAC_API int snac_correc(snac_handle *h, double dtrk);                   /* correc, src/correc.f90:7-48 */
SNAC_API int snac_updt_pressure(snac_handle *h, double alpha);           /* updt_pressure, src/updt_pressure.f90:7-47 */
SNAC_API int snac_chkdt(snac_handle *h, double *dt_cfl);                 /* chkdt, src/chkdt.f90:8-56 (global) */
SNAC_API int snac_chkdiv(snac_handle *h, double *divtot, double *divmax);/* chkdiv, src/chkdiv.f90:8-44 (global) */
/* Plane averages of the output routines on the device: the contribution of box `rank` (the sum over the boxes of a
 * block, MPI_ALLREDUCE over comm_block in the reference, and the text file stay with the host).  idir = 0..2; the
 * weights are the ghosted 1-D spacing arrays the reference passes (index lo-1:hi+1) for the two averaged directions
 * a < b (out1d) or for idir (out2d); scale = 1/(l_a l_b) resp. 1/l_idir.
 *   out1d, src/output.f90:34-133:  p1d(t)   = sum_{a,b} p(a,b,t) wa(a) wb(b) scale     -> n[idir] values
 *   out2d, src/output.f90:135-225: p2d(a,b) = sum_t p(a,b,t) w(t) scale                -> n[a]*n[b] values, a fastest */
SNAC_API int snac_out1d(snac_handle *h, int rank, int field, int idir, const double *wa, const double *wb, double scale, double *p1d);
SNAC_API int snac_out2d(snac_handle *h, int rank, int field, int idir, const double *w, double scale, double *p2d);

/* ---- coarse-grained entry point: the whole `do irk=1,3` body, src/main.f90:906-1033 -------- */
SNAC_API int snac_step(snac_handle *h, double dt);

/* ---- diagnostics ------------------------------------------------------------------------- */
/* iterations and worst relative residual of the most recent linear solve (the reference ignores
 * HYPRE's ierr and iteration counts, src/solver.f90:553-556; non-convergence is not an error). */
SNAC_API int snac_solver_stats(snac_handle *h, int *iters, double *relres);
/* V-cycles of the most recent solve averaged over the independent 2-D problems (wavenumbers): each
 * one stops on its own residual, as the reference's one-solve-per-wavenumber loop does */
SNAC_API int snac_solver_mean_iters(snac_handle *h, double *mean_iters);
/* which solver handles the pressure problem: 1 = multigrid, 2 = direct (transform + tridiagonal) */
SNAC_API int snac_plane_solver(const snac_handle *h);
/* Per-kernel-class device timing.  While enabled, every class section is bracketed by CUDA events
 * on the handle's stream and waited for, so the stream is serialised: use it to attribute time, not
 * to measure throughput.  Classes are numbered from 0; snac_profile_name returns NULL past the end. */
SNAC_API int snac_profile_enable(snac_handle *h, int on);
SNAC_API int snac_profile_get(snac_handle *h, int cls, double *ms, long long *launches);
/* grid points processed by the launches of a class (the solver kernels count only the wavenumber
 * planes that were still iterating) */
SNAC_API int snac_profile_units(snac_handle *h, int cls, double *points);
SNAC_API const char *snac_profile_name(int cls);
/* CUDA-event stopwatch on the handle's stream (start records, stop records + waits). */
SNAC_API int snac_timer_start(snac_handle *h);
SNAC_API int snac_timer_stop(snac_handle *h, double *ms);
/* number of kernels launched by this handle since creation */
SNAC_API long long snac_launch_count(const snac_handle *h);
/* blocks until all device work of the handle is complete */
SNAC_API int snac_sync(snac_handle *h);
/* CUDA stream (cudaStream_t) the handle launches on, for event timing by the caller */
SNAC_API void *snac_stream(snac_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* SNAC_B200_H */
