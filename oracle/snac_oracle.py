"""CPU oracle for SNaC's per-time-step pressure-correction pipeline.  TEST INFRASTRUCTURE ONLY.

This module is a numpy/scipy restatement of the reference's algorithm for the hot path
(reference ``src/main.f90:906-1033`` and the routines it calls).  It exists to *check* the CUDA
implementation; nothing under ``snac_b200/`` may import it.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it.

Parity pinning: the restatement is pinned by the reference's own golden vector
``tests/lid_driven_cavity/data_ldc_re1000.txt`` (64 values of v on the cavity centre line after
1500 steps, rtol 1e-7 / atol 5e-8, reference ``tests/lid_driven_cavity/test.py:14-16``) for BOTH
Poisson paths the reference test exercises (``FFT_AXIS=0`` and ``FFT_AXIS=1``,
``tests/lid_driven_cavity/testit.sh:15``) on the same 4-block decomposition: see
``tests/test_oracle_golden.py``.  The two third-party libraries on the path are absent from
``/root/reference`` and unpinned there (HYPRE: CI clones master, ``.github/workflows/compile.yml:15``;
FFTW3: system package, ``compile.yml:14``):
  * HYPRE struct solvers are restated as "the exact solution of the assembled linear system"
    (sparse LU) -- what every convergent iterative solver tends to as ``hypre_tol -> 0``;
  * FFTW r2r kinds are restated with ``scipy.fft`` (same unnormalised definitions as the FFTW
    manual: R2HC/HC2R halfcomplex, REDFT/RODFT 00/10/01/11 = DCT/DST I/II/III/IV).
At the library boundary (a transform or a single solve in isolation) parity is therefore
unpinned by the reference; end to end it is pinned by the golden vector.

Array convention: Fortran-ordered numpy arrays with one ghost layer; python index ``[i,j,k]``
corresponds to Fortran ``(lo-1+i, lo-1+j, lo-1+k)``.  ``I = slice(1,-1)`` is the interior.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft
import scipy.sparse as sp
import scipy.sparse.linalg as spla

RKCOEFF = np.array([[32.0 / 60.0, 0.0], [25.0 / 60.0, -17.0 / 60.0], [45.0 / 60.0, -25.0 / 60.0]])
I = slice(1, -1)


def _sh(a, d, s):
    """Interior-shaped view of ghosted array ``a`` shifted by ``s`` cells along direction ``d``."""
    idx = [slice(1, -1)] * 3
    n = a.shape[d]
    idx[d] = slice(1 + s, n - 1 + s)
    return a[tuple(idx)]


def _b(a1d, d, s=0):
    """Interior part of a ghosted 1-D array (shifted by s) broadcast along direction d."""
    n = a1d.shape[0]
    v = a1d[1 + s:n - 1 + s]
    shp = [1, 1, 1]
    shp[d] = v.shape[0]
    return v.reshape(shp)


# ---------------------------------------------------------------------------------------------
# momentum right-hand sides  (reference src/mom.f90)
# ---------------------------------------------------------------------------------------------
def momx_a(g, u, v, w):
    """``momx_a_vv`` (src/mom.f90:112-146)."""
    dxc, dxf, dyf, dzf = _b(g.dxc, 0), _b(g.dxf, 0), _b(g.dyf, 1), _b(g.dzf, 2)
    dxfp = _b(g.dxf, 0, 1)
    u0, uip, uim = _sh(u, 0, 0), _sh(u, 0, 1), _sh(u, 0, -1)
    uuip = 0.25 * (u0 * dyf * dzf + uip * dyf * dzf) * uip
    uuim = 0.25 * (u0 * dyf * dzf + uim * dyf * dzf) * uim
    v00, vip = _sh(v, 0, 0), _sh(v, 0, 1)
    vjm = v[1:-1, 0:-2, 1:-1]
    vipjm = v[2:, 0:-2, 1:-1]
    uvjp = 0.25 * (v00 * dxf * dzf + vip * dxfp * dzf) * _sh(u, 1, 1)
    uvjm = 0.25 * (vjm * dxf * dzf + vipjm * dxfp * dzf) * _sh(u, 1, -1)
    w00, wip = _sh(w, 0, 0), _sh(w, 0, 1)
    wkm = w[1:-1, 1:-1, 0:-2]
    wipkm = w[2:, 1:-1, 0:-2]
    uwkp = 0.25 * (w00 * dxf * dyf + wip * dxfp * dyf) * _sh(u, 2, 1)
    uwkm = 0.25 * (wkm * dxf * dyf + wipkm * dxfp * dyf) * _sh(u, 2, -1)
    return ((-uuip + uuim) + (-uvjp + uvjm) + (-uwkp + uwkm)) / (dxc * dyf * dzf)


def momy_a(g, u, v, w):
    """``momy_a_vv`` (src/mom.f90:148-183)."""
    dxf, dyc, dyf, dzf = _b(g.dxf, 0), _b(g.dyc, 1), _b(g.dyf, 1), _b(g.dzf, 2)
    dyfp = _b(g.dyf, 1, 1)
    u00, ujp = _sh(u, 0, 0), _sh(u, 1, 1)
    uim = u[0:-2, 1:-1, 1:-1]
    uimjp = u[0:-2, 2:, 1:-1]
    uvip = 0.25 * (u00 * dyf * dzf + ujp * dyfp * dzf) * _sh(v, 0, 1)
    uvim = 0.25 * (uim * dyf * dzf + uimjp * dyfp * dzf) * _sh(v, 0, -1)
    v0, vjp, vjm = _sh(v, 0, 0), _sh(v, 1, 1), _sh(v, 1, -1)
    vvjp = 0.25 * (v0 * dxf * dzf + vjp * dxf * dzf) * vjp
    vvjm = 0.25 * (v0 * dxf * dzf + vjm * dxf * dzf) * vjm
    w00, wjp = _sh(w, 0, 0), _sh(w, 1, 1)
    wkm = w[1:-1, 1:-1, 0:-2]
    wjpkm = w[1:-1, 2:, 0:-2]
    wvkp = 0.25 * (w00 * dyf * dxf + wjp * dyfp * dxf) * _sh(v, 2, 1)
    wvkm = 0.25 * (wkm * dyf * dxf + wjpkm * dyfp * dxf) * _sh(v, 2, -1)
    return ((-uvip + uvim) + (-vvjp + vvjm) + (-wvkp + wvkm)) / (dxf * dyc * dzf)


def momz_a(g, u, v, w):
    """``momz_a_vv`` (src/mom.f90:185-219)."""
    dxf, dyf, dzc, dzf = _b(g.dxf, 0), _b(g.dyf, 1), _b(g.dzc, 2), _b(g.dzf, 2)
    dzfp = _b(g.dzf, 2, 1)
    u00, ukp = _sh(u, 0, 0), _sh(u, 2, 1)
    uim = u[0:-2, 1:-1, 1:-1]
    uimkp = u[0:-2, 1:-1, 2:]
    uwip = 0.25 * (u00 * dzf * dyf + ukp * dzfp * dyf) * _sh(w, 0, 1)
    uwim = 0.25 * (uim * dzf * dyf + uimkp * dzfp * dyf) * _sh(w, 0, -1)
    v00, vkp = _sh(v, 0, 0), _sh(v, 2, 1)
    vjm = v[1:-1, 0:-2, 1:-1]
    vjmkp = v[1:-1, 0:-2, 2:]
    vwjp = 0.25 * (v00 * dzf * dxf + vkp * dzfp * dxf) * _sh(w, 1, 1)
    vwjm = 0.25 * (vjm * dzf * dxf + vjmkp * dzfp * dxf) * _sh(w, 1, -1)
    w0, wkp, wkm = _sh(w, 0, 0), _sh(w, 2, 1), _sh(w, 2, -1)
    wwkp = 0.25 * (w0 * dxf * dyf + wkp * dxf * dyf) * wkp
    wwkm = 0.25 * (w0 * dxf * dyf + wkm * dxf * dyf) * wkm
    return ((-uwip + uwim) + (-vwjp + vwjm) + (-wwkp + wwkm)) / (dxf * dyf * dzc)


def mom_d(g, q, comp, visc):
    """``mom{x,y,z}_d`` (src/mom.f90:221-327): diffusion of velocity component ``comp``."""
    out = 0.0
    for d in range(3):
        q0, qp, qm = _sh(q, 0, 0), _sh(q, d, 1), _sh(q, d, -1)
        if d == comp:   # staggered direction: gradients live at cell centres
            dp = (qp - q0) / _b(g.df(d), d, 1)
            dm = (q0 - qm) / _b(g.df(d), d, 0)
            out = out + visc * (dp - dm) / _b(g.dc(d), d)
        else:
            dp = (qp - q0) / _b(g.dc(d), d, 0)
            dm = (q0 - qm) / _b(g.dc(d), d, -1)
            out = out + visc * (dp - dm) / _b(g.df(d), d)
    return out


def mom_p(g, p, comp, bforce):
    """``mom{x,y,z}_p`` (src/mom.f90:328-386)."""
    return -(_sh(p, comp, 1) - _sh(p, comp, 0)) / _b(g.dc(comp), comp) + bforce


def scal_d(g, s, alpha):
    """``scal_d`` (src/scal.f90:89-121)."""
    out = 0.0
    for d in range(3):
        s0, sp_, sm = _sh(s, 0, 0), _sh(s, d, 1), _sh(s, d, -1)
        dp = (sp_ - s0) / _b(g.dc(d), d, 0)
        dm = (s0 - sm) / _b(g.dc(d), d, -1)
        out = out + alpha * (dp - dm) / _b(g.df(d), d)
    return out


def scal_a(g, u, v, w, s):
    """``scal_a`` (src/scal.f90:123-154)."""
    out = 0.0
    for d, vel in enumerate((u, v, w)):
        s0 = _sh(s, 0, 0)
        fm = 0.5 * (_sh(s, d, -1) + s0) * _sh(vel, d, -1)
        fp = 0.5 * (_sh(s, d, 1) + s0) * _sh(vel, d, 0)
        out = out + (-fp + fm) / _b(g.df(d), d)
    return out


# ---------------------------------------------------------------------------------------------
# boundary conditions  (reference src/bound.f90)
# ---------------------------------------------------------------------------------------------
def set_bc(ctype, ibound, idir, centered, rvalue, dr, p):
    """``set_bc`` (src/bound.f90:108-248); acts on full ghost extents of the face."""
    n = p.shape[idir] - 2

    def pl(q):
        idx = [slice(None)] * 3
        idx[idir] = q
        return tuple(idx)

    factor = rvalue
    sgn = 0.0
    if ctype == "D" and centered:
        factor = 2.0 * factor
        sgn = -1.0
    if ctype == "N":
        factor = -dr * factor if ibound == 0 else dr * factor
        sgn = 1.0
    if ctype not in ("D", "N"):
        return
    if centered:
        if ibound == 0:
            p[pl(0)] = factor + sgn * p[pl(1)]
        else:
            p[pl(n + 1)] = factor + sgn * p[pl(n)]
    elif ctype == "D":
        if ibound == 0:
            p[pl(0)] = factor
        else:
            p[pl(n)] = factor
            p[pl(n + 1)] = p[pl(n - 1)]
    elif ctype == "N":
        if ibound == 0:
            p[pl(0)] = factor + p[pl(1)]
        else:
            p[pl(n)] = factor + p[pl(n - 1)]
            p[pl(n + 1)] = p[pl(n)]


# ---------------------------------------------------------------------------------------------
# transforms  (reference src/fft.f90; FFTW r2r kinds restated with scipy.fft)
# ---------------------------------------------------------------------------------------------
def r2hc(x, axis):
    n = x.shape[axis]
    c = sfft.rfft(x, axis=axis)
    out = np.empty_like(x)
    xs = [slice(None)] * x.ndim

    def ix(s):
        t = list(xs)
        t[axis] = s
        return tuple(t)

    out[ix(slice(0, n // 2 + 1))] = c.real
    m = (n + 1) // 2 - 1  # number of stored imaginary parts
    if m > 0:
        out[ix(slice(n - 1, n - 1 - m, -1))] = c.imag[ix(slice(1, m + 1))]
    return out


def hc2r(x, axis):
    n = x.shape[axis]
    xs = [slice(None)] * x.ndim

    def ix(s):
        t = list(xs)
        t[axis] = s
        return tuple(t)

    shp = list(x.shape)
    shp[axis] = n // 2 + 1
    c = np.zeros(shp, dtype=np.complex128)
    c.real[...] = x[ix(slice(0, n // 2 + 1))]
    m = (n + 1) // 2 - 1
    if m > 0:
        c.imag[ix(slice(1, m + 1))] = x[ix(slice(n - 1, n - 1 - m, -1))]
    return sfft.irfft(c, n=n, axis=axis) * n


def find_fft(bc01, is_centered):
    """``find_fft`` (src/fft.f90:96-149): returns (fwd, bwd, norm) with fwd/bwd callables."""
    def dct(t):
        return lambda x, axis: sfft.dct(x, type=t, axis=axis)

    def dst(t):
        return lambda x, axis: sfft.dst(x, type=t, axis=axis)

    if bc01 == "FF":
        return r2hc, hc2r, (1.0, 0.0)
    if is_centered:
        return {"NN": (dct(2), dct(3), (2.0, 0.0)), "DD": (dst(2), dst(3), (2.0, 0.0)),
                "ND": (dct(4), dct(4), (2.0, 0.0)), "DN": (dst(4), dst(4), (2.0, 0.0))}[bc01]
    return {"NN": (dct(1), dct(1), (2.0, -1.0)), "DD": (dst(1), dst(1), (2.0, 1.0)),
            "ND": (dct(2), dct(3), (2.0, 0.0)), "DN": (dst(3), dst(2), (2.0, 0.0))}[bc01]


def eigenvalues(n, bc01, is_centered):
    """``eigenvalues`` (src/fft.f90:151-189). For 'DN' the reference leaves lambda undefined; the
    mathematically right values (same as 'ND') are used here."""
    l = np.arange(1, n + 1, dtype=np.float64)
    pi = np.pi
    if bc01 == "FF":
        return -4.0 * np.sin((l - 1) * pi / (1.0 * n)) ** 2
    if bc01 == "NN":
        return -4.0 * np.sin((l - 1) * pi / (2.0 * n)) ** 2
    if bc01 == "DD":
        if is_centered:
            return -4.0 * np.sin(l * pi / (2.0 * n)) ** 2
        lam = np.zeros(n)
        lam[:n - 1] = -4.0 * np.sin(l[:n - 1] * pi / (2.0 * n)) ** 2
        return lam
    if bc01 in ("ND", "DN"):
        return -4.0 * np.sin((2 * l - 1) * pi / (4.0 * n)) ** 2
    raise ValueError(bc01)


class FFTPlan:
    """``fftini`` (src/fft.f90:10-69): transform pair along ``idir`` of the interior box and
    ``normfft = 1/(norm1*(n+norm2-q))``."""

    def __init__(self, idir, n, bc01, is_centered):
        self.idir = idir
        self.q = 1 if (bc01 != "FF" and not is_centered) else 0
        self.fwd, self.bwd, norm = find_fft(bc01, is_centered)
        self.normfft = 1.0 / (norm[0] * (n + norm[1] - self.q))
        self.n = n

    def apply(self, arr_interior, direction):
        f = self.fwd if direction == 1 else self.bwd
        idx = [slice(None)] * 3
        idx[self.idir] = slice(0, self.n - self.q)
        arr_interior[tuple(idx)] = f(arr_interior[tuple(idx)], self.idir)


# ---------------------------------------------------------------------------------------------
# linear systems  (reference src/solver.f90) -- exact solves replace HYPRE
# ---------------------------------------------------------------------------------------------
def bc_factors(cbc, bc, dl, is_bound, is_centered):
    """factor/sgn tables shared by ``init_bc_rhs`` and ``init_matrix_*`` (src/solver.f90:71-94)."""
    factor = np.zeros((3, 2))
    sgn = np.zeros((3, 2))
    for q in range(3):
        for qq in (0, 1):
            if is_bound[q, qq]:
                if cbc[qq, q] == "N":
                    factor[q, qq] = 1.0 * dl[q, qq] * bc[qq, q]
                    if qq == 1:
                        factor[q, qq] = -factor[q, qq]
                    sgn[q, qq] = 1.0
                elif cbc[qq, q] == "D":
                    if is_centered[q]:
                        factor[q, qq] = -2.0 * bc[qq, q]
                        sgn[q, qq] = -1.0
                    else:
                        factor[q, qq] = -1.0 * bc[qq, q]
                        sgn[q, qq] = 0.0
    return factor, sgn


def init_bc_rhs(cbc, bc, dl, is_bound, is_centered, n, nhi, periods, d1, d2, fft_axis, bcplanes=None):
    """``init_bc_rhs`` (src/solver.f90:44-157).  ``d1[d]``/``d2[d]`` are the ghosted 1-D arrays passed
    as dx1/dx2 etc.; ``nhi[d]`` = hi-lo+1 of the (possibly shortened) range.  Returns rhs[d] with
    shape (2, n_a, n_b) (python: [ib, ...] over the two transverse interior index ranges)."""
    factor, _ = bc_factors(cbc, bc, dl, is_bound, is_centered)
    qqq = [0 if c else 1 for c in is_centered]
    tr = {0: (1, 2), 1: (0, 2), 2: (0, 1)}
    out = []
    for idir in range(3):
        a, b = tr[idir]
        r = np.zeros((2, n[a], n[b]))
        if periods[idir] == 0:
            wts = []
            for d in (a, b):
                wts.append(np.ones(n[d]) if d == fft_axis - 1 else d2[d][1:n[d] + 1])
            # weight = product of d2 over non-FFT directions, incl. the normal one
            wn = 1.0 if idir == fft_axis - 1 else None
            for ib in (0, 1):
                if not is_bound[idir, ib]:
                    continue
                i = 1 if ib == 0 else nhi[idir]          # python index of lo / hi(shortened)
                w_n = wn if wn is not None else d2[idir][i]
                weight = w_n * wts[0][:, None] * wts[1][None, :]
                val = factor[idir, ib] * weight / (d1[idir][i - (1 - ib) + qqq[idir]] * d2[idir][i])
                # transverse ranges are limited to the shortened extents as in the reference loops
                m = np.zeros_like(val)
                m[:nhi[a], :nhi[b]] = 1.0
                val = val * m
                if bcplanes is not None and bc[ib, idir] != 0.0:
                    val = val * bcplanes[idir][ib] / bc[ib, idir]
                r[ib] = val
        out.append(r)
    return out


def updt_rhs(is_bound, rhs, nhi, p):
    """``updt_rhs`` (src/bound.f90:543-585) on the (possibly shortened) interior range ``nhi``."""
    a0, a1, a2 = nhi
    if is_bound[0, 0]:
        p[1, 1:a1 + 1, 1:a2 + 1] += rhs[0][0][:a1, :a2]
    if is_bound[0, 1]:
        p[a0, 1:a1 + 1, 1:a2 + 1] += rhs[0][1][:a1, :a2]
    if is_bound[1, 0]:
        p[1:a0 + 1, 1, 1:a2 + 1] += rhs[1][0][:a0, :a2]
    if is_bound[1, 1]:
        p[1:a0 + 1, a1, 1:a2 + 1] += rhs[1][1][:a0, :a2]
    if is_bound[2, 0]:
        p[1:a0 + 1, 1:a1 + 1, 1] += rhs[2][0][:a0, :a1]
    if is_bound[2, 1]:
        p[1:a0 + 1, 1:a1 + 1, a2] += rhs[2][1][:a0, :a1]


class LinearSystem:
    """Assembled operator of ``init_matrix_3d`` / ``init_matrix_2d``+``init_n_2d_matrices``
    (src/solver.f90:158-372, 687-842, 981-1020) over *all* boxes (HYPRE's struct grid spans
    MPI_COMM_WORLD), solved exactly.

    Unknowns: interior cells ``[0:nhi)`` of every box.  ``fft_axis`` in {0,1,2,3}: for 1..3 the
    coupling along that axis is dropped and ``lambda(q)*weight`` added to the diagonal, which is
    the block-diagonal collection of the reference's per-wavenumber 2-D matrices.
    """

    def __init__(self, world, grids, cbc_of, bc_of, dl_of, is_centered, nhi_of, d1_of, d2_of,
                 fft_axis, lam=None):
        self.world, self.fft_axis = world, fft_axis
        self.nhi = [np.array(nhi_of(r), dtype=np.int64) for r in range(world.nrank)]
        self.off = np.concatenate([[0], np.cumsum([int(np.prod(n)) for n in self.nhi])])
        self.N = int(self.off[-1])
        rows, cols, vals = [], [], []
        diag_w = np.zeros(self.N)
        singular = True
        qqq = [0 if c else 1 for c in is_centered]
        for r, box in enumerate(world.boxes):
            n = self.nhi[r]
            d1, d2 = d1_of(r), d2_of(r)
            _, sgn = bc_factors(cbc_of(r), bc_of(r), dl_of(r), box.is_bound, is_centered)
            ids = (self.off[r] + np.arange(int(np.prod(n)))).reshape(tuple(n), order="F")
            w1 = [np.ones(n[d]) if d == fft_axis - 1 else d2[d][1:n[d] + 1] for d in range(3)]
            weight = w1[0][:, None, None] * w1[1][None, :, None] * w1[2][None, None, :]
            diag_w[self.off[r]:self.off[r + 1]] = weight.ravel(order="F")
            cc = np.zeros(tuple(n))
            if lam is not None and fft_axis > 0:
                shp = [1, 1, 1]
                shp[fft_axis - 1] = n[fft_axis - 1]
                cc = cc + lam[:n[fft_axis - 1]].reshape(shp) * weight
            for d in range(3):
                if d == fft_axis - 1:
                    continue
                oth = [w1[e] if e != d else np.ones(n[d]) for e in range(3)]
                wo = oth[0][:, None, None] * oth[1][None, :, None] * oth[2][None, None, :]
                shp = [1, 1, 1]
                shp[d] = n[d]
                cm = (1.0 / d1[d][0 + qqq[d]:n[d] + qqq[d]]).reshape(shp) * wo      # index i-1+qqq
                cp = (1.0 / d1[d][1 + qqq[d]:n[d] + 1 + qqq[d]]).reshape(shp) * wo  # index i+qqq
                cc = cc - (cm + cp)
                sl_lo = [slice(None)] * 3
                sl_lo[d] = 0
                sl_hi = [slice(None)] * 3
                sl_hi[d] = n[d] - 1
                sl_lo, sl_hi = tuple(sl_lo), tuple(sl_hi)
                # interior couplings
                a = [slice(None)] * 3
                b = [slice(None)] * 3
                a[d] = slice(1, None)
                b[d] = slice(0, -1)
                rows.append(ids[tuple(a)].ravel()); cols.append(ids[tuple(b)].ravel()); vals.append(cm[tuple(a)].ravel())
                rows.append(ids[tuple(b)].ravel()); cols.append(ids[tuple(a)].ravel()); vals.append(cp[tuple(b)].ravel())
                for ib, sl in ((0, sl_lo), (1, sl_hi)):
                    c = cm if ib == 0 else cp
                    nbr = int(box.nb[d, ib])
                    if box.is_bound[d, ib]:
                        if world.periods[d] == 0:
                            cc[sl] += sgn[d, ib] * c[sl]
                            if sgn[d, ib] != 1.0:
                                singular = False
                        # (is_bound with periods != 0 cannot happen: every face has a friend)
                    else:
                        nn = self.nhi[nbr]
                        nids = (self.off[nbr] + np.arange(int(np.prod(nn)))).reshape(tuple(nn), order="F")
                        s2 = [slice(None)] * 3
                        s2[d] = nn[d] - 1 if ib == 0 else 0
                        tgt = nids[tuple(s2)]
                        src = ids[sl]
                        if tgt.shape != src.shape:   # shortened staggered ranges: clip to common part
                            m = tuple(slice(0, min(x, y)) for x, y in zip(tgt.shape, src.shape))
                            tgt, src, cv = tgt[m], src[m], c[sl][m]
                        else:
                            cv = c[sl]
                        rows.append(src.ravel()); cols.append(tgt.ravel()); vals.append(cv.ravel())
            rows.append(ids.ravel()); cols.append(ids.ravel()); vals.append(cc.ravel())
        self.A0 = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                                shape=(self.N, self.N))
        self.W = diag_w
        self.lam = lam
        self.pure_neumann = singular
        self._lu = {}

    def _factor(self, alpha):
        key = float(alpha)
        if key not in self._lu:
            A = (self.A0 + sp.diags(alpha * self.W)).tocsc()
            pins = self._null_pins(alpha)
            if len(pins):
                A = A.tolil()
                for i in pins:
                    A.rows[i] = [i]
                    A.data[i] = [1.0]
                A = A.tocsc()
            self._lu[key] = (spla.splu(A), pins)
            if len(self._lu) > 8:
                self._lu.pop(next(iter(self._lu)))
        return self._lu[key]

    def _null_pins(self, alpha):
        """Rows to pin where the operator is singular (all-Neumann/periodic, alpha=0, lambda=0):
        one unknown per singular wavenumber plane."""
        if not self.pure_neumann or alpha != 0.0:
            return []
        if self.fft_axis == 0:
            return [0]
        pins = []
        n0 = self.nhi[0]
        ids = np.arange(int(np.prod(n0))).reshape(tuple(n0), order="F")
        for q in range(n0[self.fft_axis - 1]):
            if self.lam[q] == 0.0:
                sl = [0, 0, 0]
                sl[self.fft_axis - 1] = q
                pins.append(int(ids[tuple(sl)]))
        return pins

    def solve(self, rhs_list, alpha=0.0):
        """rhs_list[r]: ghosted array of box r; the interior [0:nhi) is replaced by the solution."""
        b = np.concatenate([rhs_list[r][1:n[0] + 1, 1:n[1] + 1, 1:n[2] + 1].ravel(order="F")
                            for r, n in enumerate(self.nhi)])
        lu, pins = self._factor(alpha)
        for i in pins:
            b[i] = 0.0
        x = lu.solve(b)
        for r, n in enumerate(self.nhi):
            rhs_list[r][1:n[0] + 1, 1:n[1] + 1, 1:n[2] + 1] = \
                x[self.off[r]:self.off[r + 1]].reshape(tuple(n), order="F")

    def apply(self, x_list, alpha=0.0):
        x = np.concatenate([x_list[r][1:n[0] + 1, 1:n[1] + 1, 1:n[2] + 1].ravel(order="F")
                            for r, n in enumerate(self.nhi)])
        y = self.A0 @ x + alpha * self.W * x
        return [y[self.off[r]:self.off[r + 1]].reshape(tuple(n), order="F") for r, n in enumerate(self.nhi)]


# ---------------------------------------------------------------------------------------------
# the simulation state and one RK3 step  (reference src/main.f90)
# ---------------------------------------------------------------------------------------------
class OracleSim:
    """All boxes of a run advanced together; mirrors ``program snac`` from allocation to the end of
    the time loop body (reference src/main.f90:226-1069), I/O excluded."""

    def __init__(self, params, world, grids, bgrids=None):
        self.p, self.world, self.g = params, world, grids
        self.bgrids = bgrids
        self.nr = world.nrank
        P = params
        self.fft_axis = P.fft_axis
        shp = [tuple(int(x) + 2 for x in b.n) for b in world.boxes]
        z = lambda r: np.zeros(shp[r], order="F")
        self.u = [z(r) for r in range(self.nr)]
        self.v = [z(r) for r in range(self.nr)]
        self.w = [z(r) for r in range(self.nr)]
        self.pr = [z(r) for r in range(self.nr)]
        self.pp = [z(r) for r in range(self.nr)]
        self.dudtrko = [np.zeros(tuple(b.n), order="F") for b in world.boxes]
        self.dvdtrko = [np.zeros(tuple(b.n), order="F") for b in world.boxes]
        self.dwdtrko = [np.zeros(tuple(b.n), order="F") for b in world.boxes]
        self.s = [[z(r) for r in range(self.nr)] for _ in range(P.nscal)]
        self.dsdtrko = [[np.zeros(tuple(b.n), order="F") for b in world.boxes] for _ in range(P.nscal)]
        self.blk = [P.blocks[b.block] for b in world.boxes]
        self.time, self.istep = 0.0, 0
        self.dt = 0.0
        self.alphaoi = 0.0
        self.s_alphai_o = [0.0] * P.nscal
        self.fs = np.zeros(P.nscal)
        self.u_in = None  # inflow planes: dict[(idir,ib)] -> (n,t1,t2) per box, set by set_inflow
        self.is_bound_inflow = [np.zeros((3, 2), bool) for _ in range(self.nr)]
        self._setup_solvers()

    # -- setup of solvers (src/main.f90:726-886) ------------------------------------------------
    def _setup_solvers(self):
        P, W, G = self.p, self.world, self.g
        fa = self.fft_axis
        ctr = [True, True, True]
        full = lambda r: W.boxes[r].n
        dc = lambda r: [G[r].dxc, G[r].dyc, G[r].dzc]
        df = lambda r: [G[r].dxf, G[r].dyf, G[r].dzf]
        self.rhsp = [init_bc_rhs(self.blk[r].cbcpre, self.blk[r].bcpre, G[r].dl, W.boxes[r].is_bound, ctr,
                                 W.boxes[r].n, W.boxes[r].n, W.periods, dc(r), df(r), fa)
                     for r in range(self.nr)]
        lam = None
        if fa > 0:
            d = fa - 1
            n = int(W.boxes[0].n[d])
            bc01 = self.blk[0].cbcpre[0, d] + self.blk[0].cbcpre[1, d]
            self.plan_p = FFTPlan(d, n, bc01, True)
            lam = eigenvalues(n, bc01, True) / G[0].dl[d, 0] ** 2
            self.lambda_p = lam
        self.psolver = LinearSystem(W, G, lambda r: self.blk[r].cbcpre, lambda r: self.blk[r].bcpre,
                                    lambda r: G[r].dl, ctr, full, dc, df, fa, lam)
        if P.impdiff:
            self._setup_impdiff()

    def _setup_impdiff(self):
        """Staggered Helmholtz operators for u,v,w and centred ones for scalars
        (src/main.f90:770-886)."""
        P, W, G = self.p, self.world, self.g
        fa = self.fft_axis
        self.vel_sys = []
        for comp in range(3):
            ctr = [True, True, True]
            ctr[comp] = False

            def nhi(r, comp=comp):
                n = W.boxes[r].n.copy()
                if W.boxes[r].is_bound[comp, 1]:
                    n[comp] -= 1
                return n

            def d1(r, comp=comp):
                a = [G[r].dxc, G[r].dyc, G[r].dzc]
                a[comp] = G[r].df(comp)
                return a

            def d2(r, comp=comp):
                a = [G[r].dxf, G[r].dyf, G[r].dzf]
                a[comp] = G[r].dc(comp)
                return a

            def dl(r, comp=comp):
                # dlu = [dxf_g(lo_g), dxf_g(hi_g)] in the staggered direction (src/main.f90:771-773)
                x = G[r].dl.copy()
                x[comp, 0] = self.block_df(r, comp)[1]
                x[comp, 1] = self.block_df(r, comp)[-2]
                return x

            cbc = lambda r, comp=comp: self.blk[r].cbcvel[:, :, comp]
            bc = lambda r, comp=comp: self.blk[r].bcvel[:, :, comp]
            lam = plan = None
            if fa > 0:
                d = fa - 1
                n = int(W.boxes[0].n[d])
                bc01 = self.blk[0].cbcvel[0, d, comp] + self.blk[0].cbcvel[1, d, comp]
                plan = FFTPlan(d, n, bc01, ctr[d])
                lam = eigenvalues(n, bc01, ctr[d]) / dl(0)[d, 0] ** 2
            sys_ = LinearSystem(W, G, cbc, bc, dl, ctr, nhi, d1, d2, fa, lam)
            self.vel_sys.append(dict(sys=sys_, ctr=ctr, nhi=nhi, d1=d1, d2=d2, dl=dl, cbc=cbc, bc=bc,
                                     plan=plan))
        self.scal_sys = []
        for isc in range(P.nscal):
            ctr = [True, True, True]
            full = lambda r: W.boxes[r].n
            dc = lambda r: [G[r].dxc, G[r].dyc, G[r].dzc]
            df = lambda r: [G[r].dxf, G[r].dyf, G[r].dzf]
            cbc = lambda r, isc=isc: self.blk[r].cbcscal[:, :, isc]
            bc = lambda r, isc=isc: self.blk[r].bcscal[:, :, isc]
            lam = plan = None
            if fa > 0:
                d = fa - 1
                n = int(W.boxes[0].n[d])
                bc01 = self.blk[0].cbcscal[0, d, isc] + self.blk[0].cbcscal[1, d, isc]
                plan = FFTPlan(d, n, bc01, True)
                lam = eigenvalues(n, bc01, True) / G[0].dl[d, 0] ** 2
            sys_ = LinearSystem(W, G, cbc, bc, lambda r: G[r].dl, ctr, full, dc, df, fa, lam)
            rhs = [init_bc_rhs(cbc(r), bc(r), G[r].dl, W.boxes[r].is_bound, ctr, W.boxes[r].n,
                               W.boxes[r].n, W.periods, dc(r), df(r), fa) for r in range(self.nr)]
            self.scal_sys.append(dict(sys=sys_, plan=plan, rhs=rhs, d2=df, nhi=full))

    def block_df(self, r, d):
        """Block-global cell widths (d?f_g) of box r's block, ghosted; supplied by the caller via
        ``self.bgrids`` when IMPDIFF is on."""
        return self.bgrids[self.world.boxes[r].block][d].drf

    # -- halo exchange + BCs (src/bound.f90) ---------------------------------------------------
    def updthalo(self, arrs, idir):
        """``updthalo`` (src/bound.f90:587-650) for all boxes at once; sends use pre-exchange data
        (MPI_SENDRECV semantics)."""
        W = self.world
        sends = []
        for r, box in enumerate(W.boxes):
            a = arrs[r]
            n = a.shape[idir] - 2
            lo_pl = [slice(None)] * 3
            hi_pl = [slice(None)] * 3
            lo_pl[idir] = 1
            hi_pl[idir] = n
            sends.append((a[tuple(lo_pl)].copy(), a[tuple(hi_pl)].copy()))
        for r, box in enumerate(W.boxes):
            a = arrs[r]
            n = a.shape[idir] - 2
            g_lo = [slice(None)] * 3
            g_hi = [slice(None)] * 3
            g_lo[idir] = 0
            g_hi[idir] = n + 1
            if box.nb[idir, 1] >= 0:
                a[tuple(g_hi)] = sends[int(box.nb[idir, 1])][0]
            if box.nb[idir, 0] >= 0:
                a[tuple(g_lo)] = sends[int(box.nb[idir, 0])][1]

    def bounduvw(self, is_correc):
        """``bounduvw`` (src/bound.f90:8-66)."""
        for f in (self.u, self.v, self.w):
            for idir in range(3):
                self.updthalo(f, idir)
        for r, box in enumerate(self.world.boxes):
            g, b = self.g[r], self.blk[r]
            fld = (self.u[r], self.v[r], self.w[r])
            for idir in range(3):
                n = int(box.n[idir])
                for ib in (0, 1):
                    if not box.is_bound[idir, ib]:
                        continue
                    q = 0 if ib == 0 else n   # dxc(lo-1)/dxf(lo-1) or dxc(hi)/dxf(hi)
                    for comp in range(3):
                        if comp == idir:
                            if not is_correc:
                                set_bc(b.cbcvel[ib, idir, comp], ib, idir, False, b.bcvel[ib, idir, comp],
                                       g.df(idir)[q], fld[comp])
                        else:
                            set_bc(b.cbcvel[ib, idir, comp], ib, idir, True, b.bcvel[ib, idir, comp],
                                   g.dc(idir)[q], fld[comp])

    def boundp(self, arrs, cbc_of, bc_of):
        """``boundp`` (src/bound.f90:68-106)."""
        for idir in range(3):
            self.updthalo(arrs, idir)
        for r, box in enumerate(self.world.boxes):
            g = self.g[r]
            cbc, bc = cbc_of(r), bc_of(r)
            for idir in range(3):
                n = int(box.n[idir])
                for ib in (0, 1):
                    if box.is_bound[idir, ib]:
                        q = 0 if ib == 0 else n
                        set_bc(cbc[ib, idir], ib, idir, True, bc[ib, idir], g.dc(idir)[q], arrs[r])

    def set_inflow(self, is_bound_inflow, planes):
        """planes[r][(idir,ib)] = (veln, velt1, velt2) ghosted 2-D arrays (src/main.f90:594-632)."""
        self.is_bound_inflow = is_bound_inflow
        self.u_in = planes

    def inflow(self, is_correc):
        """``inflow`` (src/bound.f90:438-477)."""
        if self.u_in is None:
            return
        for r, box in enumerate(self.world.boxes):
            fld = [self.u[r], self.v[r], self.w[r]]
            for idir in range(3):
                for ib in (0, 1):
                    if not self.is_bound_inflow[r][idir, ib]:
                        continue
                    n = int(box.n[idir])
                    qn = 0 if ib == 0 else n
                    qt = 0 if ib == 0 else n + 1
                    dq = 1 if ib == 0 else -1
                    veln, velt1, velt2 = self.u_in[r][(idir, ib)]
                    tang = [d for d in range(3) if d != idir]

                    def pl(q):
                        idx = [slice(None)] * 3
                        idx[idir] = q
                        return tuple(idx)

                    if not is_correc:
                        fld[idir][pl(qn)] = veln
                    fld[tang[0]][pl(qt)] = -fld[tang[0]][pl(qt + dq)] + 2.0 * velt1
                    fld[tang[1]][pl(qt)] = -fld[tang[1]][pl(qt + dq)] + 2.0 * velt2

    # -- pieces of the sub-step ----------------------------------------------------------------
    def rk_mom(self, irk):
        """``rk_mom`` (src/rk.f90:7-128) for all boxes."""
        P = self.p
        f1 = RKCOEFF[irk, 0] * self.dt
        f2 = RKCOEFF[irk, 1] * self.dt
        f12 = f1 + f2
        for r in range(self.nr):
            g = self.g[r]
            u, v, w, p = self.u[r], self.v[r], self.w[r], self.pr[r]
            dd = [mom_d(g, q, c, P.visc) for c, q in enumerate((u, v, w))]
            aa = [momx_a(g, u, v, w), momy_a(g, u, v, w), momz_a(g, u, v, w)]
            pg = [mom_p(g, p, c, P.bforce[c]) for c in range(3)]
            if P.boussinesq:
                s = self.s[0][r]
                buo = [-f12 * P.gacc[c] * P.beta * 0.5 * (_sh(s, c, 1) + _sh(s, c, 0)) for c in range(3)]
            old = (self.dudtrko[r], self.dvdtrko[r], self.dwdtrko[r])
            for c, q in enumerate((u, v, w)):
                if not P.impdiff:
                    rk = dd[c] + aa[c]
                    q[I, I, I] = q[I, I, I] + f1 * rk + f2 * old[c]
                else:
                    rk = 0.0 + aa[c]
                    q[I, I, I] = q[I, I, I] + f1 * rk + f2 * old[c]
                    q[I, I, I] = q[I, I, I] + f12 * dd[c]
                old[c][...] = rk
                q[I, I, I] = q[I, I, I] + f12 * pg[c]
                if P.boussinesq:
                    q[I, I, I] = q[I, I, I] + buo[c]
                if P.impdiff:
                    q[I, I, I] = q[I, I, I] - 0.5 * f12 * dd[c]

    def rk_scal(self, irk, isc):
        """``rk_scal`` (src/rk.f90:130-204) + ``bulk_forcing_s`` (src/scal.f90:181-199)."""
        P = self.p
        f1 = RKCOEFF[irk, 0] * self.dt
        f2 = RKCOEFF[irk, 1] * self.dt
        f12 = f1 + f2
        alpha = P.alphai[isc] ** (-1)
        dds = []
        for r in range(self.nr):
            g = self.g[r]
            s = self.s[isc][r]
            dd = scal_d(g, s, alpha)
            aa = scal_a(g, self.u[r], self.v[r], self.w[r], s)
            rk = aa if P.impdiff else dd + aa
            s[I, I, I] = s[I, I, I] + f1 * rk + f2 * self.dsdtrko[isc][r] + f12 * P.ssource[isc]
            if P.impdiff:
                s[I, I, I] = s[I, I, I] + f12 * dd
            self.dsdtrko[isc][r][...] = rk
            dds.append(dd)
        f = 0.0
        if P.is_sforced[isc]:
            mean = 0.0
            for r in range(self.nr):
                g = self.g[r]
                mean += np.sum(self.s[isc][r][I, I, I] * _b(g.dxf, 0) * _b(g.dyf, 1) * _b(g.dzf, 2))
            mean /= P.vol_all
            f = P.scalf[isc] - mean
        if P.impdiff:
            for r in range(self.nr):
                self.s[isc][r][I, I, I] -= 0.5 * f12 * dds[r]
        if P.is_sforced[isc]:
            for r in range(self.nr):
                self.s[isc][r][I, I, I] += f
        self.fs[isc] += f

    def fillps(self, dtrk):
        """``fillps`` (src/fillps.f90:7-54)."""
        fa = self.fft_axis
        for r in range(self.nr):
            g = self.g[r]
            u, v, w = self.u[r], self.v[r], self.w[r]
            dxf, dyf, dzf = _b(g.dxf, 0), _b(g.dyf, 1), _b(g.dzf, 2)
            du = _sh(u, 0, 0) - _sh(u, 0, -1)
            dv = _sh(v, 1, 0) - _sh(v, 1, -1)
            dw = _sh(w, 2, 0) - _sh(w, 2, -1)
            if fa == 3:
                val = dw / dzf * dxf * dyf + dv * dxf + du * dyf
            elif fa == 2:
                val = dw * dxf + dv / dyf * dxf * dzf + du * dzf
            elif fa == 1:
                val = dw * dyf + dv * dzf + du / dxf * dyf * dzf
            else:
                val = dw * dxf * dyf + dv * dxf * dzf + du * dyf * dzf
            self.pp[r][I, I, I] = val / dtrk

    def solve_poisson(self):
        """FFT build: ``fft`` -> ``solve_n_helmholtz_2d`` -> ``fft`` -> ``*normfft``
        (src/main.f90:1005-1020); else ``solve_helmholtz`` (:1022)."""
        if self.fft_axis > 0:
            for r in range(self.nr):
                self.plan_p.apply(self.pp[r][I, I, I], 1)
        self.psolver.solve(self.pp, 0.0)
        if self.fft_axis > 0:
            for r in range(self.nr):
                self.plan_p.apply(self.pp[r][I, I, I], -1)
                self.pp[r][I, I, I] *= self.plan_p.normfft

    def correc(self, dtrk):
        """``correc`` (src/correc.f90:7-48): ranges include ghost layers."""
        for r in range(self.nr):
            g = self.g[r]
            p = self.pp[r]
            self.u[r][0:-1, :, :] -= dtrk * (p[1:, :, :] - p[0:-1, :, :]) / g.dxc[0:-1, None, None]
            self.v[r][:, 0:-1, :] -= dtrk * (p[:, 1:, :] - p[:, 0:-1, :]) / g.dyc[None, 0:-1, None]
            self.w[r][:, :, 0:-1] -= dtrk * (p[:, :, 1:] - p[:, :, 0:-1]) / g.dzc[None, None, 0:-1]

    def updt_pressure(self, alpha):
        """``updt_pressure`` (src/updt_pressure.f90:7-47)."""
        for r in range(self.nr):
            g = self.g[r]
            pp = self.pp[r]
            self.pr[r][I, I, I] += pp[I, I, I]
            if self.p.impdiff:
                lap = 0.0
                for d in range(3):
                    lap = lap + ((_sh(pp, d, 1) - _sh(pp, d, 0)) / _b(g.dc(d), d, 0)
                                 - (_sh(pp, d, 0) - _sh(pp, d, -1)) / _b(g.dc(d), d, -1)) / _b(g.df(d), d)
                self.pr[r][I, I, I] += alpha * lap

    def solve_impdiff(self, sysd, fields, alphai_new, rhs, nhi_of):
        """``solve_impdiff_field_{fft,3d}`` (src/solve_helmholtz.f90:32-138)."""
        fa = self.fft_axis
        for r in range(self.nr):
            g = self.g[r]
            d2 = sysd["d2"](r)
            n = self.world.boxes[r].n
            w1 = [np.ones(n[d]) if d == fa - 1 else d2[d][1:n[d] + 1] for d in range(3)]
            fields[r][I, I, I] *= alphai_new * (w1[0][:, None, None] * w1[1][None, :, None] * w1[2][None, None, :])
            updt_rhs(self.world.boxes[r].is_bound.astype(bool), rhs[r], nhi_of(r), fields[r])
        plan = sysd["plan"]
        if fa > 0:
            for r in range(self.nr):
                plan.apply(fields[r][I, I, I], 1)
        sysd["sys"].solve(fields, alphai_new)
        if fa > 0:
            for r in range(self.nr):
                plan.apply(fields[r][I, I, I], -1)
                fields[r][I, I, I] *= plan.normfft

    # -- diagnostics ---------------------------------------------------------------------------
    def chkdt(self):
        """``chkdt`` (src/chkdt.f90:8-56)."""
        P = self.p
        dti = 0.0
        dlmin = np.inf
        for r in range(self.nr):
            g = self.g[r]
            u, v, w = self.u[r], self.v[r], self.w[r]
            dxc, dxf = _b(g.dxc, 0), _b(g.dxf, 0)
            dyc, dyf = _b(g.dyc, 1), _b(g.dyf, 1)
            dzc, dzf = _b(g.dzc, 2), _b(g.dzf, 2)
            ux = np.abs(u[I, I, I])
            vx = 0.25 * np.abs(v[I, I, I] + v[1:-1, 0:-2, 1:-1] + v[2:, 1:-1, 1:-1] + v[2:, 0:-2, 1:-1])
            wx = 0.25 * np.abs(w[I, I, I] + w[1:-1, 1:-1, 0:-2] + w[2:, 1:-1, 1:-1] + w[2:, 1:-1, 0:-2])
            dtix = ux / dxc + vx / dyf + wx / dzf
            uy = 0.25 * np.abs(u[I, I, I] + u[1:-1, 2:, 1:-1] + u[0:-2, 2:, 1:-1] + u[0:-2, 1:-1, 1:-1])
            vy = np.abs(v[I, I, I])
            wy = 0.25 * np.abs(w[I, I, I] + w[1:-1, 2:, 1:-1] + w[1:-1, 2:, 0:-2] + w[1:-1, 1:-1, 0:-2])
            dtiy = uy / dxf + vy / dyc + wy / dzf
            uz = 0.25 * np.abs(u[I, I, I] + u[0:-2, 1:-1, 1:-1] + u[0:-2, 1:-1, 2:] + u[1:-1, 1:-1, 2:])
            vz = 0.25 * np.abs(v[I, I, I] + v[1:-1, 0:-2, 1:-1] + v[1:-1, 0:-2, 2:] + v[1:-1, 1:-1, 2:])
            wz = np.abs(w[I, I, I])
            dtiz = uz / dxf + vz / dyf + wz / dzc
            dti = max(dti, float(dtix.max()), float(dtiy.max()), float(dtiz.max()))
            dlmin = min(dlmin, g.dxf.min(), g.dyf.min(), g.dzf.min())
        if dti == 0.0:
            dti = 1.0
        dtmax = np.sqrt(3.0) / dti
        if not P.impdiff:
            visc = max(P.visc, P.alpha_max)
            dtmax = min(dtmax, 1.65 / 12.0 / visc * dlmin ** 2)
        return float(dtmax)

    def chkdiv(self):
        """``chkdiv`` (src/chkdiv.f90:8-44)."""
        divtot, divmax = 0.0, 0.0
        for r in range(self.nr):
            g = self.g[r]
            u, v, w = self.u[r], self.v[r], self.w[r]
            dxf, dyf, dzf = _b(g.dxf, 0), _b(g.dyf, 1), _b(g.dzf, 2)
            div = (_sh(w, 2, 0) - _sh(w, 2, -1)) / dzf + (_sh(v, 1, 0) - _sh(v, 1, -1)) / dyf \
                + (_sh(u, 0, 0) - _sh(u, 0, -1)) / dxf
            divmax = max(divmax, float(np.abs(div).max()))
            divtot += float(np.sum(div * dxf * dyf * dzf / self.p.vol_all))
        return divtot, divmax

    def set_dt(self):
        P = self.p
        self.dt_cfl = self.chkdt()
        self.dt = P.dt_f if P.dt_f > 0.0 else min(P.cfl * self.dt_cfl, P.dtmax)
        return self.dt

    # -- initial BCs (src/main.f90:562-568,633-642) ---------------------------------------------
    def init_bc(self):
        self.bounduvw(False)
        self.boundp(self.pr, lambda r: self.blk[r].cbcpre, lambda r: self.blk[r].bcpre)
        for isc in range(self.p.nscal):
            self.boundp(self.s[isc], lambda r, isc=isc: self.blk[r].cbcscal[:, :, isc],
                        lambda r, isc=isc: self.blk[r].bcscal[:, :, isc])
        self.inflow(False)

    # -- one time step (src/main.f90:897-1033) -------------------------------------------------
    def step(self):
        P = self.p
        self.istep += 1
        self.time += self.dt
        self.fs[:] = 0.0
        for irk in range(3):
            self.substep(irk)
        if P.icheck > 0 and self.istep % max(P.icheck, 1) == 0:
            self.set_dt()
            self.divtot, self.divmax = self.chkdiv()

    def substep(self, irk):
        P = self.p
        dtrk = float(RKCOEFF[irk, 0] + RKCOEFF[irk, 1]) * self.dt
        for isc in range(P.nscal):
            self.rk_scal(irk, isc)
            if P.impdiff:
                alpha = -(P.alphai[isc] ** (-1)) * dtrk / 2.0
                alphai = alpha ** (-1)
                sd = self.scal_sys[isc]
                self.solve_impdiff(sd, self.s[isc], alphai, sd["rhs"], sd["nhi"])
                self.s_alphai_o[isc] = alphai
            self.boundp(self.s[isc], lambda r, isc=isc: self.blk[r].cbcscal[:, :, isc],
                        lambda r, isc=isc: self.blk[r].bcscal[:, :, isc])
        alpha = -P.visc * dtrk / 2.0
        self.rk_mom(irk)
        if P.impdiff:
            alphai = alpha ** (-1)
            for comp, fld in enumerate((self.u, self.v, self.w)):
                sd = self.vel_sys[comp]
                bcpl = None
                if self.u_in is not None:
                    bcpl = self._impdiff_bc_planes(comp)
                rhs = [init_bc_rhs(sd["cbc"](r), sd["bc"](r), sd["dl"](r), self.world.boxes[r].is_bound,
                                   sd["ctr"], self.world.boxes[r].n, sd["nhi"](r), self.world.periods,
                                   sd["d1"](r), sd["d2"](r), self.fft_axis,
                                   None if bcpl is None else bcpl[r]) for r in range(self.nr)]
                self.solve_impdiff(sd, fld, alphai, rhs, sd["nhi"])
            self.alphaoi = alphai
        self.bounduvw(False)
        self.inflow(False)
        self.fillps(dtrk)
        for r in range(self.nr):
            updt_rhs(self.world.boxes[r].is_bound, self.rhsp[r], self.world.boxes[r].n, self.pp[r])
        self.solve_poisson()
        self.boundp(self.pp, lambda r: self.blk[r].cbcpre, lambda r: self.blk[r].bcpre)
        self.correc(dtrk)
        self.bounduvw(True)
        self.inflow(True)
        self.updt_pressure(alpha)
        self.boundp(self.pr, lambda r: self.blk[r].cbcpre, lambda r: self.blk[r].bcpre)

    def _impdiff_bc_planes(self, comp):
        return None
