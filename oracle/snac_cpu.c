/* CPU restatement (C + OpenMP) of SNaC's per-time-step pressure-correction pipeline for the
 * benchmark workload: one triperiodic block on a uniform grid with FFT_AXIS = 3 (the reference's
 * examples/_manuscript_taylor_green_vortex).  TEST / BASELINE INFRASTRUCTURE ONLY: it is the
 * "cpu_baseline" and "--impl reference" arm of bench.py (kind "port": the reference itself cannot be
 * built here -- no Fortran compiler, MPI, HYPRE or FFTW in the image) and nothing under snac_b200/
 * links or calls it.  Parity pinning: tests/test_cpu_port.py checks it against the numpy oracle
 * (oracle/snac_oracle.py, itself pinned by the reference's golden vector) on the same inputs.
 *
 * It follows the reference's own sweep structure, one OpenMP-parallel loop nest per Fortran routine
 * (no fusion -- that is what the reference executes on the CPU):
 *   rk_mom            src/rk.f90:7-128      (momx/y/z_d src/mom.f90:221-327, _a_vv :112-219, _p :328-386)
 *   bounduvw/boundp   src/bound.f90:8-106   (periodic halo copies; no physical faces in this case)
 *   fillps            src/fillps.f90:7-54
 *   fft               src/fft.f90:85-94     (FFTW R2HC / HC2R, halfcomplex order; here two real pencils
 *                                            per complex radix-2 transform)
 *   solve_n_helmholtz_2d src/solver.f90:843-922 (one 2-D solve per wavenumber; HYPRE PFMG is restated
 *                                            as a geometric multigrid V(1,1), red-black Gauss-Seidel, same
 *                                            stopping rule: relative residual < tol, initial guess = po)
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
  int n1, n2, n3;
  double dx, dy, dz, visc;
  double *u, *v, *w, *p, *pp, *po, *dudtrko, *dvdtrko, *dwdtrko, *dudtrk, *dvdtrk, *dwdtrk;
  double *lam;          /* eigenvalues / dz^2 per halfcomplex index */
  double tol;
  int maxiter;
  long mg_cycles;       /* statistics: V-cycles summed over planes and solves */
  long mg_solves;
  /* multigrid hierarchy (per thread scratch) */
  int nlev, *ln1, *ln2;
} Sim;

#define IDX(S, i, j, k) ((size_t)(i) + (size_t)((S)->n1 + 2) * ((size_t)(j) + (size_t)((S)->n2 + 2) * (size_t)(k)))

static const double RKC[3][2] = {{32.0 / 60.0, 0.0}, {25.0 / 60.0, -17.0 / 60.0}, {45.0 / 60.0, -25.0 / 60.0}};

int snac_cpu_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- halos: periodic in all three directions, x then y then z (src/bound.f90:587-650) ------------- */
static void halo(const Sim *S, double *a) {
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 0; k <= n3 + 1; ++k)
    for (int j = 0; j <= n2 + 1; ++j) {
      a[IDX(S, 0, j, k)] = a[IDX(S, n1, j, k)];
      a[IDX(S, n1 + 1, j, k)] = a[IDX(S, 1, j, k)];
    }
#pragma omp parallel for schedule(static)
  for (int k = 0; k <= n3 + 1; ++k)
    for (int i = 0; i <= n1 + 1; ++i) {
      a[IDX(S, i, 0, k)] = a[IDX(S, i, n2, k)];
      a[IDX(S, i, n2 + 1, k)] = a[IDX(S, i, 1, k)];
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j <= n2 + 1; ++j)
    for (int i = 0; i <= n1 + 1; ++i) {
      a[IDX(S, i, j, 0)] = a[IDX(S, i, j, n3)];
      a[IDX(S, i, j, n3 + 1)] = a[IDX(S, i, j, 1)];
    }
}

/* ---- momentum right-hand sides, one sweep per routine as in src/mom.f90 (uniform spacings) -------- */
static void mom_d(const Sim *S, const double *q, double *dq) {   /* momx_d / momy_d / momz_d */
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  const double cx = S->visc / (S->dx * S->dx), cy = S->visc / (S->dy * S->dy), cz = S->visc / (S->dz * S->dz);
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
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  const double *u = S->u, *v = S->v, *w = S->w;
  const double dx = S->dx, dy = S->dy, dz = S->dz;
  const double rvol = 1.0 / (dx * dy * dz);
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const size_t c = IDX(S, i, j, k);
        double fxp, fxm, fyp, fym, fzp, fzm;
        if (comp == 0) {
          const double ayz = dy * dz, axz = dx * dz, axy = dx * dy;
          const double uip = u[IDX(S, i + 1, j, k)], uim = u[IDX(S, i - 1, j, k)];
          fxp = 0.25 * (u[c] * ayz + uip * ayz) * uip;
          fxm = 0.25 * (u[c] * ayz + uim * ayz) * uim;
          fyp = 0.25 * (v[c] * axz + v[IDX(S, i + 1, j, k)] * axz) * u[IDX(S, i, j + 1, k)];
          fym = 0.25 * (v[IDX(S, i, j - 1, k)] * axz + v[IDX(S, i + 1, j - 1, k)] * axz) * u[IDX(S, i, j - 1, k)];
          fzp = 0.25 * (w[c] * axy + w[IDX(S, i + 1, j, k)] * axy) * u[IDX(S, i, j, k + 1)];
          fzm = 0.25 * (w[IDX(S, i, j, k - 1)] * axy + w[IDX(S, i + 1, j, k - 1)] * axy) * u[IDX(S, i, j, k - 1)];
        } else if (comp == 1) {
          const double ayz = dy * dz, axz = dx * dz, axy = dx * dy;
          const double vjp = v[IDX(S, i, j + 1, k)], vjm = v[IDX(S, i, j - 1, k)];
          fxp = 0.25 * (u[c] * ayz + u[IDX(S, i, j + 1, k)] * ayz) * v[IDX(S, i + 1, j, k)];
          fxm = 0.25 * (u[IDX(S, i - 1, j, k)] * ayz + u[IDX(S, i - 1, j + 1, k)] * ayz) * v[IDX(S, i - 1, j, k)];
          fyp = 0.25 * (v[c] * axz + vjp * axz) * vjp;
          fym = 0.25 * (v[c] * axz + vjm * axz) * vjm;
          fzp = 0.25 * (w[c] * axy + w[IDX(S, i, j + 1, k)] * axy) * v[IDX(S, i, j, k + 1)];
          fzm = 0.25 * (w[IDX(S, i, j, k - 1)] * axy + w[IDX(S, i, j + 1, k - 1)] * axy) * v[IDX(S, i, j, k - 1)];
        } else {
          const double ayz = dy * dz, axz = dx * dz, axy = dx * dy;
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

static void mom_p(const Sim *S, int comp, double *dq) {   /* pressure gradient (bforce = 0 here) */
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  const double *p = S->p;
  const double rd = comp == 0 ? 1.0 / S->dx : (comp == 1 ? 1.0 / S->dy : 1.0 / S->dz);
  const int di = comp == 0, dj = comp == 1, dk = comp == 2;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) dq[IDX(S, i, j, k)] = -(p[IDX(S, i + di, j + dj, k + dk)] - p[IDX(S, i, j, k)]) * rd;
}

static void zero_int(const Sim *S, double *a) {
  const size_t tot = (size_t)(S->n1 + 2) * (S->n2 + 2) * (S->n3 + 2);
#pragma omp parallel for schedule(static)
  for (size_t t = 0; t < tot; ++t) a[t] = 0.0;
}

/* rk_mom (src/rk.f90:38-111): u += f1*dudtrk + f2*dudtrko + f12*dpdl; dudtrko = dudtrk */
static void rk_mom(Sim *S, int irk, double dt) {
  const double f1 = RKC[irk][0] * dt, f2 = RKC[irk][1] * dt, f12 = f1 + f2;
  double *q[3] = {S->u, S->v, S->w}, *dq[3] = {S->dudtrk, S->dvdtrk, S->dwdtrk}, *dqo[3] = {S->dudtrko, S->dvdtrko, S->dwdtrko};
  const size_t tot = (size_t)(S->n1 + 2) * (S->n2 + 2) * (S->n3 + 2);
  double *dpdl = S->pp;   /* scratch: pp is rebuilt by fillps afterwards */
  for (int c = 0; c < 3; ++c) { zero_int(S, dq[c]); mom_d(S, q[c], dq[c]); mom_a(S, c, dq[c]); }
  for (int c = 0; c < 3; ++c) {
    mom_p(S, c, dpdl);
    double *qq = q[c], *d = dq[c], *d_o = dqo[c];
    const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= n3; ++k)
      for (int j = 1; j <= n2; ++j)
        for (int i = 1; i <= n1; ++i) {
          const size_t x = IDX(S, i, j, k);
          qq[x] += f1 * d[x] + f2 * d_o[x] + f12 * dpdl[x];
        }
  }
  for (int c = 0; c < 3; ++c) {   /* swap history */
    double *t = dq[c];
    if (c == 0) { S->dudtrk = S->dudtrko; S->dudtrko = t; }
    if (c == 1) { S->dvdtrk = S->dvdtrko; S->dvdtrko = t; }
    if (c == 2) { S->dwdtrk = S->dwdtrko; S->dwdtrko = t; }
  }
  (void)tot;
}

/* fillps (src/fillps.f90:32-50, FFT build along z: divergence times the cell area normal to z, over dtrk) */
static void fillps(const Sim *S, double dtrk) {
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  const double dti = 1.0 / dtrk, dx = S->dx, dy = S->dy, rdz = 1.0 / S->dz;
  const double *u = S->u, *v = S->v, *w = S->w;
  double *pp = S->pp;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const size_t c = IDX(S, i, j, k);
        pp[c] = ((w[c] - w[IDX(S, i, j, k - 1)]) * rdz * dx * dy + (v[c] - v[IDX(S, i, j - 1, k)]) * dx +
                 (u[c] - u[IDX(S, i - 1, j, k)]) * dy) * dti;
      }
}

/* ---- transforms: FFTW R2HC / HC2R along z, two real pencils per complex radix-2 FFT ---------------- */
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

/* dir = +1: pp (physical, interior) -> halfcomplex coefficients in place; dir = -1: back, times 1/n */
static void fft_z(const Sim *S, double *a, int dir) {
  const int n1 = S->n1, n2 = S->n2, n = S->n3;
  const long npen = (long)n1 * n2;
#pragma omp parallel
  {
    double *re = (double *)malloc(sizeof(double) * n), *im = (double *)malloc(sizeof(double) * n);
#pragma omp for schedule(static)
    for (long pp2 = 0; pp2 < (npen + 1) / 2; ++pp2) {
      const long pa = 2 * pp2, pb = pa + 1 < npen ? pa + 1 : pa;
      const int ia = (int)(pa % n1) + 1, ja = (int)(pa / n1) + 1, ib = (int)(pb % n1) + 1, jb = (int)(pb / n1) + 1;
      if (dir > 0) {
        for (int k = 0; k < n; ++k) { re[k] = a[IDX(S, ia, ja, k + 1)]; im[k] = a[IDX(S, ib, jb, k + 1)]; }
        cfft(re, im, n, -1);
        /* A[k] = (Z[k] + conj Z[n-k])/2, B[k] = (Z[k] - conj Z[n-k])/(2i); halfcomplex: r0..r_{n/2}, i_{n/2-1}..i_1 */
        for (int k = 0; k <= n / 2; ++k) {
          const int m = (n - k) % n;
          const double ar = 0.5 * (re[k] + re[m]), ai = 0.5 * (im[k] - im[m]);
          const double br = 0.5 * (im[k] + im[m]), bi = -0.5 * (re[k] - re[m]);
          a[IDX(S, ia, ja, k + 1)] = ar;
          if (k > 0 && k < n - k) a[IDX(S, ia, ja, n - k + 1)] = ai;
          if (pb != pa) {
            a[IDX(S, ib, jb, k + 1)] = br;
            if (k > 0 && k < n - k) a[IDX(S, ib, jb, n - k + 1)] = bi;
          }
        }
        /* (the writes above only touch indices k and n-k, which were read into re/im before) */
      } else {
        /* Z[k] = A[k] + i B[k] with the Hermitian extension of both */
        for (int k = 0; k <= n / 2; ++k) {
          const double ar = a[IDX(S, ia, ja, k + 1)], br = a[IDX(S, ib, jb, k + 1)];
          const double ai = (k > 0 && k < n - k) ? a[IDX(S, ia, ja, n - k + 1)] : 0.0;
          const double bi = (k > 0 && k < n - k) ? a[IDX(S, ib, jb, n - k + 1)] : 0.0;
          re[k] = ar - bi; im[k] = ai + br;
          if (k > 0 && k < n - k) { re[n - k] = ar + bi; im[n - k] = -ai + br; }
        }
        cfft(re, im, n, +1);
        const double sc = 1.0 / n;
        for (int k = 0; k < n; ++k) {
          a[IDX(S, ia, ja, k + 1)] = re[k] * sc;
          if (pb != pa) a[IDX(S, ib, jb, k + 1)] = im[k] * sc;
        }
      }
    }
    free(re); free(im);
  }
}

/* ---- one 2-D solve per wavenumber: geometric multigrid on the periodic uniform x-y plane ---------- */
/* operator (matrix of init_matrix_2d + lambda*area, src/solver.f90:779-832,1014):                     */
/*   (dy/dx)(xL + xR) + (dx/dy)(xD + xU) + (-2 dy/dx - 2 dx/dy + lam dx dy) x = b                       */
typedef struct { int n1, n2; double cx, cy, area; double *x, *b; } Lev;

static void mg_smooth(const Lev *L, double lam, int color) {
  const int n1 = L->n1, n2 = L->n2;
  const double diag = -2.0 * L->cx - 2.0 * L->cy + lam * L->area;
  if (diag == 0.0) return;
  const double rd = 1.0 / diag;
  for (int j = 0; j < n2; ++j) {
    const int jm = j ? j - 1 : n2 - 1, jp = j < n2 - 1 ? j + 1 : 0;
    for (int i = (color + j) & 1; i < n1; i += 2) {
      const int im = i ? i - 1 : n1 - 1, ip = i < n1 - 1 ? i + 1 : 0;
      L->x[i + n1 * j] = (L->b[i + n1 * j] - L->cx * (L->x[im + n1 * j] + L->x[ip + n1 * j]) -
                          L->cy * (L->x[i + n1 * jm] + L->x[i + n1 * jp])) * rd;
    }
  }
}
static double mg_resid(const Lev *L, double lam, int i, int j) {
  const int n1 = L->n1, n2 = L->n2;
  const int im = i ? i - 1 : n1 - 1, ip = i < n1 - 1 ? i + 1 : 0, jm = j ? j - 1 : n2 - 1, jp = j < n2 - 1 ? j + 1 : 0;
  const double diag = -2.0 * L->cx - 2.0 * L->cy + lam * L->area;
  return L->b[i + n1 * j] - (L->cx * (L->x[im + n1 * j] + L->x[ip + n1 * j]) + L->cy * (L->x[i + n1 * jm] + L->x[i + n1 * jp]) +
                             diag * L->x[i + n1 * j]);
}
static void mg_vcycle(Lev *lev, int nlev, int l, double lam) {
  Lev *F = &lev[l];
  if (l == nlev - 1) {
    const int ns = F->n1 * F->n2 < 20 ? F->n1 * F->n2 : 20;
    for (int s = 0; s < ns; ++s) { mg_smooth(F, lam, 0); mg_smooth(F, lam, 1); }
    return;
  }
  mg_smooth(F, lam, 0); mg_smooth(F, lam, 1);
  Lev *C = &lev[l + 1];
  for (int J = 0; J < C->n2; ++J)
    for (int I = 0; I < C->n1; ++I) {
      C->b[I + C->n1 * J] = mg_resid(F, lam, 2 * I, 2 * J) + mg_resid(F, lam, 2 * I + 1, 2 * J) + mg_resid(F, lam, 2 * I, 2 * J + 1) +
                            mg_resid(F, lam, 2 * I + 1, 2 * J + 1);
      C->x[I + C->n1 * J] = 0.0;
    }
  mg_vcycle(lev, nlev, l + 1, lam);
  for (int j = 0; j < F->n2; ++j)
    for (int i = 0; i < F->n1; ++i) F->x[i + F->n1 * j] += C->x[(i >> 1) + C->n1 * (j >> 1)];
  mg_smooth(F, lam, 1); mg_smooth(F, lam, 0);
}

static void solve_planes(Sim *S) {
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  long cycles = 0;
#pragma omp parallel reduction(+ : cycles)
  {
    /* per-thread hierarchy */
    Lev lev[32];
    int nlev = 0;
    {
      int a = n1, b = n2;
      double dx = S->dx, dy = S->dy;
      for (;;) {
        Lev *L = &lev[nlev++];
        L->n1 = a; L->n2 = b; L->cx = dy / dx; L->cy = dx / dy; L->area = dx * dy;
        L->x = (double *)malloc(sizeof(double) * a * b);
        L->b = (double *)malloc(sizeof(double) * a * b);
        if ((a & 1) || (b & 1) || a < 4 || b < 4 || nlev >= 32) break;
        a >>= 1; b >>= 1; dx *= 2.0; dy *= 2.0;
      }
    }
#pragma omp for schedule(dynamic, 1)
    for (int k = 1; k <= n3; ++k) {
      Lev *T = &lev[0];
      const double lam = S->lam[k - 1];
      double b2 = 0.0;
      for (int j = 0; j < n2; ++j)
        for (int i = 0; i < n1; ++i) {
          const double bb = S->pp[IDX(S, i + 1, j + 1, k)];
          T->b[i + n1 * j] = bb;
          T->x[i + n1 * j] = S->po[IDX(S, i + 1, j + 1, k)];   /* initial guess: previous solution (src/solver.f90:870) */
          b2 += bb * bb;
        }
      if (b2 == 0.0) {
        for (int t = 0; t < n1 * n2; ++t) T->x[t] = 0.0;   /* HYPRE returns x = 0 for b = 0 */
      } else {
        for (int it = 0; it < S->maxiter; ++it) {
          mg_vcycle(lev, nlev, 0, lam);
          ++cycles;
          double r2 = 0.0;
          for (int j = 0; j < n2; ++j)
            for (int i = 0; i < n1; ++i) { const double r = mg_resid(T, lam, i, j); r2 += r * r; }
          if (r2 < S->tol * S->tol * b2) break;
        }
      }
      for (int j = 0; j < n2; ++j)
        for (int i = 0; i < n1; ++i) {
          const double x = T->x[i + n1 * j];
          S->pp[IDX(S, i + 1, j + 1, k)] = x;
          S->po[IDX(S, i + 1, j + 1, k)] = x;
        }
    }
    for (int l = 0; l < nlev; ++l) { free(lev[l].x); free(lev[l].b); }
  }
  S->mg_cycles += cycles;
  S->mg_solves += n3;
}

/* correc (src/correc.f90:23-47; interior and ghost ranges as in the reference) */
static void correc(const Sim *S, double dtrk) {
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  const double fx = dtrk / S->dx, fy = dtrk / S->dy, fz = dtrk / S->dz;
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
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
#pragma omp parallel for collapse(2) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) S->p[IDX(S, i, j, k)] += S->pp[IDX(S, i, j, k)];
}

/* ---- public interface (ctypes) ------------------------------------------------------------------- */
void *snac_cpu_create(int n1, int n2, int n3, double lx, double ly, double lz, double visc, double tol, int maxiter) {
  Sim *S = (Sim *)calloc(1, sizeof(Sim));
  S->n1 = n1; S->n2 = n2; S->n3 = n3;
  S->dx = lx / n1; S->dy = ly / n2; S->dz = lz / n3;
  S->visc = visc; S->tol = tol; S->maxiter = maxiter;
  const size_t tot = (size_t)(n1 + 2) * (n2 + 2) * (n3 + 2);
  double **f[] = {&S->u, &S->v, &S->w, &S->p, &S->pp, &S->po, &S->dudtrko, &S->dvdtrko, &S->dwdtrko, &S->dudtrk, &S->dvdtrk, &S->dwdtrk};
  for (size_t q = 0; q < sizeof(f) / sizeof(f[0]); ++q) {
    *f[q] = (double *)malloc(sizeof(double) * tot);
    zero_int(S, *f[q]);   /* first touch in parallel */
  }
  S->lam = (double *)malloc(sizeof(double) * n3);
  for (int l = 0; l < n3; ++l) {   /* src/fft.f90:151-189 ('FF'), divided by dl^2 (src/solver.f90:1137) */
    const double s = sin(l * M_PI / n3);
    S->lam[l] = -4.0 * s * s / (S->dz * S->dz);
  }
  return S;
}
void snac_cpu_destroy(void *h) {
  Sim *S = (Sim *)h;
  double *f[] = {S->u, S->v, S->w, S->p, S->pp, S->po, S->dudtrko, S->dvdtrko, S->dwdtrko, S->dudtrk, S->dvdtrk, S->dwdtrk, S->lam};
  for (size_t q = 0; q < sizeof(f) / sizeof(f[0]); ++q) free(f[q]);
  free(S);
}
/* which: 0 u, 1 v, 2 w, 3 p; ghosted Fortran-ordered arrays */
void snac_cpu_set(void *h, int which, const double *a) {
  Sim *S = (Sim *)h;
  double *d = which == 0 ? S->u : (which == 1 ? S->v : (which == 2 ? S->w : S->p));
  memcpy(d, a, sizeof(double) * (size_t)(S->n1 + 2) * (S->n2 + 2) * (S->n3 + 2));
}
void snac_cpu_get(void *h, int which, double *a) {
  Sim *S = (Sim *)h;
  const double *d = which == 0 ? S->u : (which == 1 ? S->v : (which == 2 ? S->w : S->p));
  memcpy(a, d, sizeof(double) * (size_t)(S->n1 + 2) * (S->n2 + 2) * (S->n3 + 2));
}
/* the `do irk = 1,3` body of src/main.f90:906-1033 */
void snac_cpu_step(void *h, double dt) {
  Sim *S = (Sim *)h;
  for (int irk = 0; irk < 3; ++irk) {
    const double dtrk = (RKC[irk][0] + RKC[irk][1]) * dt;
    rk_mom(S, irk, dt);
    halo(S, S->u); halo(S, S->v); halo(S, S->w);     /* bounduvw(.false.) */
    fillps(S, dtrk);
    fft_z(S, S->pp, +1);
    solve_planes(S);
    fft_z(S, S->pp, -1);
    halo(S, S->pp);                                   /* boundp(pp) */
    correc(S, dtrk);
    halo(S, S->u); halo(S, S->v); halo(S, S->w);     /* bounduvw(.true.) */
    updt_pressure(S);
    halo(S, S->p);                                    /* boundp(p) */
  }
}
/* chkdt (src/chkdt.f90:29-56): returns dt_cfl */
double snac_cpu_chkdt(void *h) {
  Sim *S = (Sim *)h;
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  const double *u = S->u, *v = S->v, *w = S->w;
  double dti = 0.0;
#pragma omp parallel for collapse(2) reduction(max : dti) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const double ux = fabs(u[IDX(S, i, j, k)]);
        const double vx = 0.25 * fabs(v[IDX(S, i, j, k)] + v[IDX(S, i, j - 1, k)] + v[IDX(S, i + 1, j, k)] + v[IDX(S, i + 1, j - 1, k)]);
        const double wx = 0.25 * fabs(w[IDX(S, i, j, k)] + w[IDX(S, i, j, k - 1)] + w[IDX(S, i + 1, j, k)] + w[IDX(S, i + 1, j, k - 1)]);
        const double dtix = ux / S->dx + vx / S->dy + wx / S->dz;
        const double uy = 0.25 * fabs(u[IDX(S, i, j, k)] + u[IDX(S, i, j + 1, k)] + u[IDX(S, i - 1, j + 1, k)] + u[IDX(S, i - 1, j, k)]);
        const double vy = fabs(v[IDX(S, i, j, k)]);
        const double wy = 0.25 * fabs(w[IDX(S, i, j, k)] + w[IDX(S, i, j + 1, k)] + w[IDX(S, i, j + 1, k - 1)] + w[IDX(S, i, j, k - 1)]);
        const double dtiy = uy / S->dx + vy / S->dy + wy / S->dz;
        const double uz = 0.25 * fabs(u[IDX(S, i, j, k)] + u[IDX(S, i - 1, j, k)] + u[IDX(S, i - 1, j, k + 1)] + u[IDX(S, i, j, k + 1)]);
        const double vz = 0.25 * fabs(v[IDX(S, i, j, k)] + v[IDX(S, i, j - 1, k)] + v[IDX(S, i, j - 1, k + 1)] + v[IDX(S, i, j, k + 1)]);
        const double wz = fabs(w[IDX(S, i, j, k)]);
        const double dtiz = uz / S->dx + vz / S->dy + wz / S->dz;
        const double m = fmax(dtix, fmax(dtiy, dtiz));
        if (m > dti) dti = m;
      }
  if (dti == 0.0) dti = 1.0;
  double dlmin = fmin(S->dx, fmin(S->dy, S->dz));
  const double dtv = 1.65 / 12.0 / S->visc * dlmin * dlmin;
  return fmin(sqrt(3.0) / dti, dtv);
}
double snac_cpu_divmax(void *h) {
  Sim *S = (Sim *)h;
  const int n1 = S->n1, n2 = S->n2, n3 = S->n3;
  double dm = 0.0;
#pragma omp parallel for collapse(2) reduction(max : dm) schedule(static)
  for (int k = 1; k <= n3; ++k)
    for (int j = 1; j <= n2; ++j)
      for (int i = 1; i <= n1; ++i) {
        const double d = (S->w[IDX(S, i, j, k)] - S->w[IDX(S, i, j, k - 1)]) / S->dz + (S->v[IDX(S, i, j, k)] - S->v[IDX(S, i, j - 1, k)]) / S->dy +
                         (S->u[IDX(S, i, j, k)] - S->u[IDX(S, i - 1, j, k)]) / S->dx;
        if (fabs(d) > dm) dm = fabs(d);
      }
  return dm;
}
double snac_cpu_mean_cycles(void *h) {
  Sim *S = (Sim *)h;
  return S->mg_solves ? (double)S->mg_cycles / (double)S->mg_solves : 0.0;
}
