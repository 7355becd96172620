/* CPU restatement (C + OpenMP) of SNaC's per-time-step pressure-correction pipeline for the
 * benchmark workloads: one rectangular domain on a uniform grid, every direction either periodic
 * ('F' faces) or bounded by two no-slip walls (velocity 'D' = 0, pressure 'N' = 0), transforms along
 * one periodic direction (FFT_AXIS = 1, 2 or 3).  That covers the reference's
 * examples/_manuscript_taylor_green_vortex (triperiodic, FFT_AXIS = 3), _manuscript_turbulent_channel
 * (walls in z, FFT_AXIS = 2) and _manuscript_turbulent_duct / periodic_duct (walls in y and z,
 * FFT_AXIS = 1; the friend blocks of the duct form one rectangle, and the discrete problem does not
 * depend on how it is cut into blocks).
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY: it is the "cpu_baseline", the "--impl reference" arm and the
 * `check` of bench.py (kind "port": the reference itself cannot be built here -- no Fortran compiler,
 * MPI, HYPRE or FFTW in the image) and nothing under snac_b200/ links or calls it.  Parity pinning:
 * tests/test_cpu_port.py checks it against the numpy oracle (oracle/snac_oracle.py, itself pinned by
 * the reference's golden vector) on the same inputs for all three topologies.
 *
 * It follows the reference's own sweep structure, one OpenMP-parallel loop nest per Fortran routine
 * (no fusion -- that is what the reference executes on the CPU):
 *   rk_mom            src/rk.f90:7-128      (momx/y/z_d src/mom.f90:221-327, _a_vv :112-219, _p :328-386)
 *   bounduvw/boundp   src/bound.f90:8-106   (updthalo :587-650 as periodic copies, set_bc :108-248)
 *   fillps            src/fillps.f90:7-54
 *   fft               src/fft.f90:85-94     (FFTW R2HC / HC2R, halfcomplex order; here two real pencils
 *                                            per complex radix-2 transform)
 *   solve_n_helmholtz_2d src/solver.f90:843-922 (one 2-D solve per wavenumber, matrix of init_matrix_2d
 *                                            :687-842; HYPRE PFMG is restated as a geometric multigrid
 *                                            V(1,1), red-black Gauss-Seidel, semi-coarsening on anisotropic
 *                                            planes as PFMG does, same stopping rule: relative residual
 *                                            < tol, initial guess = po)
 *   correc            src/correc.f90:7-48
 *   updt_pressure     src/updt_pressure.f90:7-47
 *   chkdt / chkdiv    src/chkdt.f90:8-56, src/chkdiv.f90:8-44
 * Arrays are Fortran-ordered with one ghost layer: a(i,j,k) at a[i + (n1+2)*(j + (n2+2)*k)], i = 0..n1+1.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int n[3];
  double dl[3], visc, bforce[3];
  int wall[3];          /* 0: periodic, 1: two no-slip walls */
  int fa;               /* transform direction (0-based) */
  double *u, *v, *w, *p, *pp, *po, *dudtrko, *dvdtrko, *dwdtrko, *dudtrk, *dvdtrk, *dwdtrk;
  double *lam;          /* eigenvalues / dl^2 per halfcomplex index */
  double tol;
  int maxiter;
  long mg_cycles;       /* statistics: V-cycles summed over planes and solves */
  long mg_solves;
  size_t st[3];         /* strides of the ghosted arrays */
} Sim;

#define IDX(S, i, j, k) ((size_t)(i) + (S)->st[1] * (size_t)(j) + (S)->st[2] * (size_t)(k))

static const double RKC[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};

int snac_cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void snac_cpu_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- halos and boundary conditions, x then y then z over the full ghost extents ------------------- */
/* (updthalo src/bound.f90:587-650 for the periodic directions, set_bc src/bound.f90:108-248 on walls) */
/* kind: 0 centred Dirichlet 0 (ghost = -interior), 1 centred Neumann 0 (ghost = interior),             */
/*       2 staggered Dirichlet 0 (q(0) = 0; q(n) = 0, q(n+1) = q(n-1)), 3 nothing (is_correc)           */
static void bound_dir(const Sim *S, double *a, int d, int kind) {
  const int n = S->n[d];
  const int da = d == 0 ? 1 : 0, db = d == 2 ? 1 : 2;
  const int na = S->n[da] + 2, nb = S->n[db] + 2;
  const size_t sd = S->st[d], sa = S->st[da], sb = S->st[db];
  if (S->wall[d] && kind == 3) return;
#pragma omp parallel for schedule(static)
  for (int ib = 0; ib < nb; ++ib)
    for (int ia = 0; ia < na; ++ia) {
      double *q = a + sa * (size_t)ia + sb * (size_t)ib;
      if (!S->wall[d]) {
        q[0] = q[sd * (size_t)n];
        q[sd * (size_t)(n + 1)] = q[sd];
      } else if (kind == 0) {
        q[0] = -q[sd];
        q[sd * (size_t)(n + 1)] = -q[sd * (size_t)n];
      } else if (kind == 1) {
        q[0] = q[sd];
        q[sd * (size_t)(n + 1)] = q[sd * (size_t)n];
      } else {
        q[0] = 0.0;
        q[sd * (size_t)n] = 0.0;
        q[sd * (size_t)(n + 1)] = q[sd * (size_t)(n - 1)];
      }
    }
}
/* bounduvw (src/bound.f90:8-66): is_correc skips the wall-normal component */
static void bounduvw(const Sim *S, int is_correc) {
  double *q[3] = {S->u, S->v, S->w};
  /* the reference does all halos first (x,y,z per field) and then the physical faces; with periodic
     copies and wall fills acting on different directions of a rectangular domain the result is the
     same as direction by direction as long as the order x, y, z is kept */
  for (int d = 0; d < 3; ++d)
    if (!S->wall[d]) for (int c = 0; c < 3; ++c) bound_dir(S, q[c], d, 0);
  for (int d = 0; d < 3; ++d)
    if (S->wall[d]) for (int c = 0; c < 3; ++c) bound_dir(S, q[c], d, c == d ? (is_correc ? 3 : 2) : 0);
}
static void boundp(const Sim *S, double *p) {
  for (int d = 0; d < 3; ++d) if (!S->wall[d]) bound_dir(S, p, d, 1);
  for (int d = 0; d < 3; ++d) if (S->wall[d]) bound_dir(S, p, d, 1);
}

/* ---- momentum right-hand sides, one sweep per routine as in src/mom.f90 (uniform spacings) -------- */
static void mom_d(const Sim *S, const double *q, double *dq) {   /* momx_d / momy_d / momz_d */
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  const double cx = S->visc / (S->dl[0] * S->dl[0]), cy = S->visc / (S->dl[1] * S->dl[1]), cz = S->visc / (S->dl[2] * S->dl[2]);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const size_t c = IDX(S, i, j, k);
        dq[c] += cx * ((q[IDX(S, i + 1, j, k)] - q[c]) - (q[c] - q[IDX(S, i - 1, j, k)])) +
                 cy * ((q[IDX(S, i, j + 1, k)] - q[c]) - (q[c] - q[IDX(S, i, j - 1, k)])) +
                 cz * ((q[IDX(S, i, j, k + 1)] - q[c]) - (q[c] - q[IDX(S, i, j, k - 1)]));
      }
}

/* advection in the volume-weighted form of mom?_a_vv (src/mom.f90:112-219); comp = 0,1,2 */
static void mom_a(const Sim *S, int comp, double *dq) {
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  const double *u = S->u, *v = S->v, *w = S->w;
  const double dx = S->dl[0], dy = S->dl[1], dz = S->dl[2];
  const double rvol = 1.0 / (dx * dy * dz);
  const double ayz = dy * dz, axz = dx * dz, axy = dx * dy;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const size_t c = IDX(S, i, j, k);
        double fxp, fxm, fyp, fym, fzp, fzm;
        if (comp == 0) {
          const double uip = u[IDX(S, i + 1, j, k)], uim = u[IDX(S, i - 1, j, k)];
          fxp = 0.25 * (u[c] * ayz + uip * ayz) * uip;
          fxm = 0.25 * (u[c] * ayz + uim * ayz) * uim;
          fyp = 0.25 * (v[c] * axz + v[IDX(S, i + 1, j, k)] * axz) * u[IDX(S, i, j + 1, k)];
          fym = 0.25 * (v[IDX(S, i, j - 1, k)] * axz + v[IDX(S, i + 1, j - 1, k)] * axz) * u[IDX(S, i, j - 1, k)];
          fzp = 0.25 * (w[c] * axy + w[IDX(S, i + 1, j, k)] * axy) * u[IDX(S, i, j, k + 1)];
          fzm = 0.25 * (w[IDX(S, i, j, k - 1)] * axy + w[IDX(S, i + 1, j, k - 1)] * axy) * u[IDX(S, i, j, k - 1)];
        } else if (comp == 1) {
          const double vjp = v[IDX(S, i, j + 1, k)], vjm = v[IDX(S, i, j - 1, k)];
          fxp = 0.25 * (u[c] * ayz + u[IDX(S, i, j + 1, k)] * ayz) * v[IDX(S, i + 1, j, k)];
          fxm = 0.25 * (u[IDX(S, i - 1, j, k)] * ayz + u[IDX(S, i - 1, j + 1, k)] * ayz) * v[IDX(S, i - 1, j, k)];
          fyp = 0.25 * (v[c] * axz + vjp * axz) * vjp;
          fym = 0.25 * (v[c] * axz + vjm * axz) * vjm;
          fzp = 0.25 * (w[c] * axy + w[IDX(S, i, j + 1, k)] * axy) * v[IDX(S, i, j, k + 1)];
          fzm = 0.25 * (w[IDX(S, i, j, k - 1)] * axy + w[IDX(S, i, j + 1, k - 1)] * axy) * v[IDX(S, i, j, k - 1)];
        } else {
          const double wkp = w[IDX(S, i, j, k + 1)], wkm = w[IDX(S, i, j, k - 1)];
          fxp = 0.25 * (u[c] * ayz + u[IDX(S, i, j, k + 1)] * ayz) * w[IDX(S, i + 1, j, k)];
          fxm = 0.25 * (u[IDX(S, i - 1, j, k)] * ayz + u[IDX(S, i - 1, j, k + 1)] * ayz) * w[IDX(S, i - 1, j, k)];
          fyp = 0.25 * (v[c] * axz + v[IDX(S, i, j, k + 1)] * axz) * w[IDX(S, i, j + 1, k)];
          fym = 0.25 * (v[IDX(S, i, j - 1, k)] * axz + v[IDX(S, i, j - 1, k + 1)] * axz) * w[IDX(S, i, j - 1, k)];
          fzp = 0.25 * (w[c] * axy + wkp * axy) * wkp;
          fzm = 0.25 * (w[c] * axy + wkm * axy) * wkm;
        }
        dq[c] += (-fxp + fxm - fyp + fym - fzp + fzm) * rvol;
      }
}

static void mom_p(const Sim *S, int comp, double *dq) {   /* pressure gradient + body force (src/mom.f90:342,361,380) */
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  const double *p = S->p;
  const double rd = 1.0 / S->dl[comp], bf = S->bforce[comp];
  const int di = comp == 0, dj = comp == 1, dk = comp == 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) dq[IDX(S, i, j, k)] = -(p[IDX(S, i + di, j + dj, k + dk)] - p[IDX(S, i, j, k)]) * rd + bf;
}

static void zero_all(const Sim *S, double *a) {
  const size_t tot = S->st[2] * (size_t)(S->n[2] + 2);
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < tot; ++t) a[t] = 0.0;
}

/* rk_mom (src/rk.f90:38-111): u += f1*dudtrk + f2*dudtrko + f12*dpdl; dudtrko = dudtrk */
static void rk_mom(Sim *S, int irk, double dt) {
  const double f1 = RKC[irk][0] * dt, f2 = RKC[irk][1] * dt, f12 = f1 + f2;
  double *q[3] = {S->u, S->v, S->w}, *dq[3] = {S->dudtrk, S->dvdtrk, S->dwdtrk}, *dqo[3] = {S->dudtrko, S->dvdtrko, S->dwdtrko};
  double *dpdl = S->pp;   /* scratch: pp is rebuilt by fillps afterwards */
  for (int c = 0; c < 3; ++c) { zero_all(S, dq[c]); mom_d(S, q[c], dq[c]); mom_a(S, c, dq[c]); }
  for (int c = 0; c < 3; ++c) {
    mom_p(S, c, dpdl);
    double *qq = q[c], *d = dq[c], *d_o = dqo[c];
    const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= n3; ++k)
      for (int j = 1; j <= n2; ++j)
        for (int i = 1; i <= n1; ++i) {
          const size_t x = IDX(S, i, j, k);
          qq[x] += f1 * d[x] + f2 * d_o[x] + f12 * dpdl[x];
        }
  }
  double *t;   /* swap the histories */
  t = S->dudtrk; S->dudtrk = S->dudtrko; S->dudtrko = t;
  t = S->dvdtrk; S->dvdtrk = S->dvdtrko; S->dvdtrko = t;
  t = S->dwdtrk; S->dwdtrk = S->dwdtrko; S->dwdtrko = t;
}

/* fillps (src/fillps.f90:32-50, FFT build: the divergence times the cell area normal to the transform axis, over dtrk) */
static void fillps(const Sim *S, double dtrk) {
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  double wq[3], gq[3];
  for (int d = 0; d < 3; ++d) { wq[d] = d == S->fa ? 1.0 : S->dl[d]; gq[d] = d == S->fa ? 1.0 / S->dl[d] : 1.0; }
  const double cu = gq[0] * wq[1] * wq[2] / dtrk, cv = gq[1] * wq[0] * wq[2] / dtrk, cw = gq[2] * wq[0] * wq[1] / dtrk;
  const double *u = S->u, *v = S->v, *w = S->w;
  double *pp = S->pp;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const size_t c = IDX(S, i, j, k);
        pp[c] = (w[c] - w[IDX(S, i, j, k - 1)]) * cw + (v[c] - v[IDX(S, i, j - 1, k)]) * cv + (u[c] - u[IDX(S, i - 1, j, k)]) * cu;
      }
}

/* ---- transforms: FFTW R2HC / HC2R along the transform axis, two real pencils per complex radix-2 FFT */
static void cfft(double *re, double *im, int n, int sign) {   /* in place, natural order in and out */
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
  }
  for (int len = 2; len <= n; len <<= 1) {
    const double ang = sign * 2.0 * M_PI / len;
    const double wr = cos(ang), wi = sin(ang);
    for (int i = 0; i < n; i += len) {
      double cr = 1.0, ci = 0.0;
      for (int k = 0; k < len / 2; ++k) {
        const int a = i + k, b = a + len / 2;
        const double xr = re[b] * cr - im[b] * ci, xi = re[b] * ci + im[b] * cr;
        re[b] = re[a] - xr; im[b] = im[a] - xi;
        re[a] += xr; im[a] += xi;
        const double t = cr * wr - ci * wi;
        ci = cr * wi + ci * wr;
        cr = t;
      }
    }
  }
}

/* dir = +1: interior of a (physical) -> halfcomplex coefficients in place; dir = -1: back, times 1/n */
static void fft_axis(const Sim *S, double *a, int dir) {
  const int fa = S->fa, n = S->n[fa];
  const int da = fa == 0 ? 1 : 0, db = fa == 2 ? 1 : 2;
  const int na = S->n[da];
  const long npen = (long)na * S->n[db];
  const size_t sf = S->st[fa], sa = S->st[da], sb = S->st[db];
#pragma omp parallel
  {
    double *re = (double *)malloc(sizeof(double) * n), *im = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(static)
    for (long pp2 = 0; pp2 < (npen + 1) / 2; ++pp2) {
      const long pa = 2 * pp2, pb = pa + 1 < npen ? pa + 1 : pa;
      double *A = a + sa * (size_t)(pa % na + 1) + sb * (size_t)(pa / na + 1) + sf;   /* element 1 of the pencil */
      double *B = a + sa * (size_t)(pb % na + 1) + sb * (size_t)(pb / na + 1) + sf;
      if (dir > 0) {
        for (int k = 0; k < n; ++k) { re[k] = A[sf * (size_t)k]; im[k] = B[sf * (size_t)k]; }
        cfft(re, im, n, -1);
        /* A[k] = (Z[k] + conj Z[n-k])/2, B[k] = (Z[k] - conj Z[n-k])/(2i); halfcomplex: r0..r_{n/2}, i_{n/2-1}..i_1 */
        for (int k = 0; k <= n / 2; ++k) {
          const int m = (n - k) % n;
          const double ar = 0.5 * (re[k] + re[m]), ai = 0.5 * (im[k] - im[m]);
          const double br = 0.5 * (im[k] + im[m]), bi = -0.5 * (re[k] - re[m]);
          A[sf * (size_t)k] = ar;
          if (k > 0 && k < n - k) A[sf * (size_t)(n - k)] = ai;
          if (pb != pa) {
            B[sf * (size_t)k] = br;
            if (k > 0 && k < n - k) B[sf * (size_t)(n - k)] = bi;
          }
        }
      } else {
        /* Z[k] = A[k] + i B[k] with the Hermitian extension of both */
        for (int k = 0; k <= n / 2; ++k) {
          const double ar = A[sf * (size_t)k], br = B[sf * (size_t)k];
          const double ai = (k > 0 && k < n - k) ? A[sf * (size_t)(n - k)] : 0.0;
          const double bi = (k > 0 && k < n - k) ? B[sf * (size_t)(n - k)] : 0.0;
          re[k] = ar - bi; im[k] = ai + br;
          if (k > 0 && k < n - k) { re[n - k] = ar + bi; im[n - k] = -ai + br; }
        }
        cfft(re, im, n, +1);
        const double sc = 1.0 / n;
        for (int k = 0; k < n; ++k) {
          A[sf * (size_t)k] = re[k] * sc;
          if (pb != pa) B[sf * (size_t)k] = im[k] * sc;
        }
      }
    }
    free(re); free(im);
  }
}

/* ---- one 2-D solve per wavenumber: geometric multigrid on the uniform plane of the two other directions */
/* operator (matrix of init_matrix_2d + lambda*area, src/solver.f90:779-832,1014), directions a < b:        */
/*   ca (xL + xR) + cb (xD + xU) + (-2 ca - 2 cb + lam area) x = b,   ca = db/da, cb = da/db, area = da db  */
/* a Neumann wall drops the coupling across it and adds it to the diagonal (src/solver.f90:788-823)         */
typedef struct { int n1, n2, c1, c2; double ca, cb, area, h1, h2; double *x, *b; } Lev;

static inline double mg_diag(const Lev *L, const int wall[2], int i, int j, double lam) {
  double d = -2.0 * L->ca - 2.0 * L->cb + lam * L->area;
  if (wall[0] && (i == 0)) d += L->ca;
  if (wall[0] && (i == L->n1 - 1)) d += L->ca;
  if (wall[1] && (j == 0)) d += L->cb;
  if (wall[1] && (j == L->n2 - 1)) d += L->cb;
  return d;
}
static inline double mg_off(const Lev *L, const int wall[2], int i, int j) {
  const int n1 = L->n1, n2 = L->n2;
  double s = 0.0;
  if (i > 0) s += L->ca * L->x[i - 1 + n1 * j]; else if (!wall[0]) s += L->ca * L->x[n1 - 1 + n1 * j];
  if (i < n1 - 1) s += L->ca * L->x[i + 1 + n1 * j]; else if (!wall[0]) s += L->ca * L->x[n1 * j];
  if (j > 0) s += L->cb * L->x[i + n1 * (j - 1)]; else if (!wall[1]) s += L->cb * L->x[i + n1 * (n2 - 1)];
  if (j < n2 - 1) s += L->cb * L->x[i + n1 * (j + 1)]; else if (!wall[1]) s += L->cb * L->x[i];
  return s;
}
static void mg_smooth(const Lev *L, const int wall[2], double lam, int color) {
  const int n1 = L->n1, n2 = L->n2;
  for (int j = 0; j < n2; ++j)
    for (int i = (color + j) & 1; i < n1; i += 2) {
      const double d = mg_diag(L, wall, i, j, lam);
      if (d != 0.0) L->x[i + n1 * j] = (L->b[i + n1 * j] - mg_off(L, wall, i, j)) / d;
    }
}
static inline double mg_resid(const Lev *L, const int wall[2], double lam, int i, int j) {
  return L->b[i + L->n1 * j] - (mg_off(L, wall, i, j) + mg_diag(L, wall, i, j, lam) * L->x[i + L->n1 * j]);
}
static void mg_vcycle(Lev *lev, const int wall[2], int nlev, int l, double lam) {
  Lev *F = &lev[l];
  if (l == nlev - 1) {
    const int ns = F->n1 * F->n2 < 20 ? F->n1 * F->n2 : 20;
    for (int s = 0; s < ns; ++s) { mg_smooth(F, wall, lam, 0); mg_smooth(F, wall, lam, 1); }
    return;
  }
  mg_smooth(F, wall, lam, 0); mg_smooth(F, wall, lam, 1);
  Lev *C = &lev[l + 1];
  for (int J = 0; J < C->n2; ++J)
    for (int I = 0; I < C->n1; ++I) {
      double s = 0.0;
      for (int dj = 0; dj <= F->c2; ++dj)
        for (int di = 0; di <= F->c1; ++di) s += mg_resid(F, wall, lam, (F->c1 ? 2 * I : I) + di, (F->c2 ? 2 * J : J) + dj);
      C->b[I + C->n1 * J] = s;
      C->x[I + C->n1 * J] = 0.0;
    }
  mg_vcycle(lev, wall, nlev, l + 1, lam);
  for (int j = 0; j < F->n2; ++j)
    for (int i = 0; i < F->n1; ++i) F->x[i + F->n1 * j] += C->x[(F->c1 ? i >> 1 : i) + C->n1 * (F->c2 ? j >> 1 : j)];
  mg_smooth(F, wall, lam, 1); mg_smooth(F, wall, lam, 0);
}

static void solve_planes(Sim *S) {
  const int fa = S->fa, nq = S->n[fa];
  const int da = fa == 0 ? 1 : 0, db = fa == 2 ? 1 : 2;
  const int n1 = S->n[da], n2 = S->n[db];
  const size_t sf = S->st[fa], sa = S->st[da], sb = S->st[db];
  const int wall[2] = {S->wall[da], S->wall[db]};
  long cycles = 0;
#pragma omp parallel reduction(+ : cycles)
  {
    /* per-thread hierarchy: a direction is coarsened while its mesh width is within 1.5x of the finer one */
    Lev lev[40];
    int nlev = 0;
    {
      int a = n1, b = n2;
      double h1 = S->dl[da], h2 = S->dl[db];
      for (;;) {
        Lev *L = &lev[nlev++];
        L->n1 = a; L->n2 = b; L->h1 = h1; L->h2 = h2;
        L->ca = h2 / h1; L->cb = h1 / h2; L->area = h1 * h2;
        L->x = (double *)malloc(sizeof(double) * a * b);
        L->b = (double *)malloc(sizeof(double) * a * b);
        const int can1 = !(a & 1) && a >= 4, can2 = !(b & 1) && b >= 4;
        const double hmin = can1 && can2 ? fmin(h1, h2) : (can1 ? h1 : h2);
        L->c1 = can1 && h1 < 1.5 * hmin;
        L->c2 = can2 && h2 < 1.5 * hmin;
        if ((!L->c1 && !L->c2) || nlev >= 40) { L->c1 = L->c2 = 0; break; }
        if (L->c1) { a >>= 1; h1 *= 2.0; }
        if (L->c2) { b >>= 1; h2 *= 2.0; }
      }
    }
#pragma omp for schedule(dynamic, 1)
    for (int q = 1; q <= nq; ++q) {
      Lev *T = &lev[0];
      const double lam = S->lam[q - 1];
      double b2 = 0.0;
      for (int j = 0; j < n2; ++j)
        for (int i = 0; i < n1; ++i) {
          const size_t c = sf * (size_t)q + sa * (size_t)(i + 1) + sb * (size_t)(j + 1);
          const double bb = S->pp[c];
          T->b[i + n1 * j] = bb;
          T->x[i + n1 * j] = S->po[c];   /* initial guess: previous solution (src/solver.f90:870) */
          b2 += bb * bb;
        }
      if (b2 == 0.0) {
        for (int t = 0; t < n1 * n2; ++t) T->x[t] = 0.0;   /* HYPRE returns x = 0 for b = 0 */
      } else {
        for (int it = 0; it < S->maxiter; ++it) {
          mg_vcycle(lev, wall, nlev, 0, lam);
          ++cycles;
          double r2 = 0.0;
          for (int j = 0; j < n2; ++j)
            for (int i = 0; i < n1; ++i) { const double r = mg_resid(T, wall, lam, i, j); r2 += r * r; }
          if (r2 < S->tol * S->tol * b2) break;
        }
      }
      for (int j = 0; j < n2; ++j)
        for (int i = 0; i < n1; ++i) {
          const size_t c = sf * (size_t)q + sa * (size_t)(i + 1) + sb * (size_t)(j + 1);
          const double x = T->x[i + n1 * j];
          S->pp[c] = x;
          S->po[c] = x;
        }
    }
    for (int l = 0; l < nlev; ++l) { free(lev[l].x); free(lev[l].b); }
  }
  S->mg_cycles += cycles;
  S->mg_solves += nq;
}

/* correc (src/correc.f90:23-47): the normal component from lo-1 (on a wall that face value stays 0: pp(0) = pp(1)) */
static void correc(const Sim *S, double dtrk) {
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  const double fx = dtrk / S->dl[0], fy = dtrk / S->dl[1], fz = dtrk / S->dl[2];
  const double *pp = S->pp;
  double *u = S->u, *v = S->v, *w = S->w;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 0; i <= n1; ++i) u[IDX(S, i, j, k)] -= fx * (pp[IDX(S, i + 1, j, k)] - pp[IDX(S, i, j, k)]);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 0; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) v[IDX(S, i, j, k)] -= fy * (pp[IDX(S, i, j + 1, k)] - pp[IDX(S, i, j, k)]);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) w[IDX(S, i, j, k)] -= fz * (pp[IDX(S, i, j, k + 1)] - pp[IDX(S, i, j, k)]);
}

static void updt_pressure(const Sim *S) {
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) S->p[IDX(S, i, j, k)] += S->pp[IDX(S, i, j, k)];
}

/* ---- public interface (ctypes) ------------------------------------------------------------------- */
/* walls[d] = 1: no-slip walls in direction d (else periodic); fft_axis = 1,2,3 (a periodic direction) */
void *snac_cpu_create(const int *n, const double *lengths, const int *walls, int fft_axis, double visc, const double *bforce,
                      double tol, int maxiter) {
  if (fft_axis < 1 || fft_axis > 3 || walls[fft_axis - 1]) return NULL;
  Sim *S = (Sim *)calloc(1, sizeof(Sim));
  for (int d = 0; d < 3; ++d) {
    S->n[d] = n[d];
    S->dl[d] = lengths[d] / n[d];
    S->wall[d] = walls[d];
    S->bforce[d] = bforce ? bforce[d] : 0.0;
  }
  S->fa = fft_axis - 1;
  S->st[0] = 1; S->st[1] = (size_t)(n[0] + 2); S->st[2] = (size_t)(n[0] + 2) * (size_t)(n[1] + 2);
  S->visc = visc; S->tol = tol; S->maxiter = maxiter;
  const size_t tot = S->st[2] * (size_t)(n[2] + 2);
  double **f[] = {&S->u, &S->v, &S->w, &S->p, &S->pp, &S->po, &S->dudtrko, &S->dvdtrko, &S->dwdtrko, &S->dudtrk, &S->dvdtrk, &S->dwdtrk};
  for (size_t q = 0; q < sizeof(f) / sizeof(f[0]); ++q) {
    *f[q] = (double *)malloc(sizeof(double) * tot);
    if (!*f[q]) return NULL;
    zero_all(S, *f[q]);   /* first touch in parallel */
  }
  const int nq = n[S->fa];
  S->lam = (double *)malloc(sizeof(double) * nq);
  for (int l = 0; l < nq; ++l) {   /* src/fft.f90:151-189 ('FF'), divided by dl^2 (src/solver.f90:1137) */
    const double s = sin(l * M_PI / nq);
    S->lam[l] = -4.0 * s * s / (S->dl[S->fa] * S->dl[S->fa]);
  }
  return S;
}
void snac_cpu_destroy(void *h) {
  Sim *S = (Sim *)h;
  double *f[] = {S->u, S->v, S->w, S->p, S->pp, S->po, S->dudtrko, S->dvdtrko, S->dwdtrko, S->dudtrk, S->dvdtrk, S->dwdtrk, S->lam};
  for (size_t q = 0; q < sizeof(f) / sizeof(f[0]); ++q) free(f[q]);
  free(S);
}
static double *field_of(Sim *S, int which) { return which == 0 ? S->u : (which == 1 ? S->v : (which == 2 ? S->w : S->p)); }
/* which: 0 u, 1 v, 2 w, 3 p; ghosted Fortran-ordered arrays */
void snac_cpu_set(void *h, int which, const double *a) {
  Sim *S = (Sim *)h;
  memcpy(field_of(S, which), a, sizeof(double) * S->st[2] * (size_t)(S->n[2] + 2));
}
void snac_cpu_get(void *h, int which, double *a) {
  Sim *S = (Sim *)h;
  memcpy(a, field_of(S, which), sizeof(double) * S->st[2] * (size_t)(S->n[2] + 2));
}
/* bounduvw(.false.) + boundp(p) on the current fields (src/main.f90:562-568) */
void snac_cpu_init_bc(void *h) {
  Sim *S = (Sim *)h;
  bounduvw(S, 0);
  boundp(S, S->p);
}
/* the `do irk = 1,3` body of src/main.f90:906-1033 */
void snac_cpu_step(void *h, double dt) {
  Sim *S = (Sim *)h;
  for (int irk = 0; irk < 3; ++irk) {
    const double dtrk = (RKC[irk][0] + RKC[irk][1]) * dt;
    rk_mom(S, irk, dt);
    bounduvw(S, 0);
    fillps(S, dtrk);
    fft_axis(S, S->pp, +1);
    solve_planes(S);
    fft_axis(S, S->pp, -1);
    boundp(S, S->pp);
    correc(S, dtrk);
    bounduvw(S, 1);
    updt_pressure(S);
    boundp(S, S->p);
  }
}
/* chkdt (src/chkdt.f90:29-56): returns dt_cfl */
double snac_cpu_chkdt(void *h) {
  Sim *S = (Sim *)h;
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  const double *u = S->u, *v = S->v, *w = S->w;
  const double rdx = 1.0 / S->dl[0], rdy = 1.0 / S->dl[1], rdz = 1.0 / S->dl[2];
  double dti = 0.0;
#pragma omp parallel for collapse(2) reduction(max : dti) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const double ux = fabs(u[IDX(S, i, j, k)]);
        const double vx = 0.25 * fabs(v[IDX(S, i, j, k)] + v[IDX(S, i, j - 1, k)] + v[IDX(S, i + 1, j, k)] + v[IDX(S, i + 1, j - 1, k)]);
        const double wx = 0.25 * fabs(w[IDX(S, i, j, k)] + w[IDX(S, i, j, k - 1)] + w[IDX(S, i + 1, j, k)] + w[IDX(S, i + 1, j, k - 1)]);
        const double dtix = ux * rdx + vx * rdy + wx * rdz;
        const double uy = 0.25 * fabs(u[IDX(S, i, j, k)] + u[IDX(S, i, j + 1, k)] + u[IDX(S, i - 1, j + 1, k)] + u[IDX(S, i - 1, j, k)]);
        const double vy = fabs(v[IDX(S, i, j, k)]);
        const double wy = 0.25 * fabs(w[IDX(S, i, j, k)] + w[IDX(S, i, j + 1, k)] + w[IDX(S, i, j + 1, k - 1)] + w[IDX(S, i, j, k - 1)]);
        const double dtiy = uy * rdx + vy * rdy + wy * rdz;
        const double uz = 0.25 * fabs(u[IDX(S, i, j, k)] + u[IDX(S, i - 1, j, k)] + u[IDX(S, i - 1, j, k + 1)] + u[IDX(S, i, j, k + 1)]);
        const double vz = 0.25 * fabs(v[IDX(S, i, j, k)] + v[IDX(S, i, j - 1, k)] + v[IDX(S, i, j - 1, k + 1)] + v[IDX(S, i, j, k + 1)]);
        const double wz = fabs(w[IDX(S, i, j, k)]);
        const double dtiz = uz * rdx + vz * rdy + wz * rdz;
        const double m = fmax(dtix, fmax(dtiy, dtiz));
        if (m > dti) dti = m;
      }
  if (dti == 0.0) dti = 1.0;
  const double dlmin = fmin(S->dl[0], fmin(S->dl[1], S->dl[2]));
  const double dtv = 1.65 / 12.0 / S->visc * dlmin * dlmin;
  return fmin(sqrt(3.0) / dti, dtv);
}
double snac_cpu_divmax(void *h) {
  Sim *S = (Sim *)h;
  const int n1 = S->n[0], n2 = S->n[1], n3 = S->n[2];
  double dm = 0.0;
#pragma omp parallel for collapse(2) reduction(max : dm) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const double d = (S->w[IDX(S, i, j, k)] - S->w[IDX(S, i, j, k - 1)]) / S->dl[2] +
                         (S->v[IDX(S, i, j, k)] - S->v[IDX(S, i, j - 1, k)]) / S->dl[1] +
                         (S->u[IDX(S, i, j, k)] - S->u[IDX(S, i - 1, j, k)]) / S->dl[0];
        if (fabs(d) > dm) dm = fabs(d);
      }
  return dm;
}
double snac_cpu_mean_cycles(void *h) {
  Sim *S = (Sim *)h;
  return S->mg_solves ? (double)S->mg_cycles / (double)S->mg_solves : 0.0;
}
