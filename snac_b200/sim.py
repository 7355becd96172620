"""Host-side driver over the C ABI: the twin of ``program snac``'s time loop.

Mirrors the reference driver (``src/main.f90:226-253`` allocation, ``:562-568`` initial boundary
conditions, ``:660-662`` first ``dt``, ``:893-1069`` time loop) with every field resident on the
device.  The method names are the reference's subroutine names and take the same arguments; all of
them are thin calls into ``libsnacb200.so`` -- no numerics happen in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, List, Optional

import numpy as np

from . import lib as L
from .case import Case

RKCOEFF = np.array([[32.0 / 60.0, 0.0], [25.0 / 60.0, -17.0 / 60.0], [45.0 / 60.0, -25.0 / 60.0]])


def _fill2(dst, src):       # src[ib, idir] (python/Fortran-like) -> dst[idir][ib]
    for d in range(3):
        for ib in range(2):
            dst[d][ib] = src[ib, d]


class Simulation:
    def __init__(self, case: Case, ranks: Optional[Iterable[int]] = None, device: int = 0,
                 nproc: int = 1, myproc: int = 0, rank_owner: Optional[List[int]] = None,
                 tol: Optional[float] = None, maxiter: Optional[int] = None,
                 mg_npre: int = 0, mg_npost: int = 0, mg_nbottom: int = 0, mg_generic: bool = False,
                 solver: str = "auto"):
        self.case = case
        P, W = case.params, case.world
        self.lib = L.load_library()
        self.nrank = W.nrank
        owner = list(rank_owner) if rank_owner is not None else [0] * W.nrank
        self.ranks = list(ranks) if ranks is not None else [r for r in range(W.nrank) if owner[r] == myproc]
        prm = L.snac_params()
        prm.abi_version = L.SNAC_ABI_VERSION
        prm.device = device
        prm.nrank, prm.nproc, prm.myproc = W.nrank, nproc, myproc
        self._owner = (C.c_int * W.nrank)(*owner)
        prm.rank_owner = C.cast(self._owner, C.POINTER(C.c_int))
        for d in range(3):
            prm.periods[d] = int(W.periods[d])
            prm.bforce[d] = float(P.bforce[d])
            prm.gacc[d] = float(P.gacc[d])
        prm.fft_axis, prm.impdiff, prm.boussinesq, prm.nscal = int(P.fft_axis), int(P.impdiff), int(P.boussinesq), int(P.nscal)
        prm.visc, prm.beta, prm.vol_all = float(P.visc), float(P.beta), float(P.vol_all)
        for i in range(P.nscal):
            prm.alphai[i] = float(P.alphai[i])
            prm.ssource[i] = float(P.ssource[i])
            prm.is_sforced[i] = int(P.is_sforced[i])
            prm.scalf[i] = float(P.scalf[i])
        prm.hypre_tol = float(P.hypre_tol if tol is None else tol)
        prm.hypre_maxiter = int(P.hypre_maxiter if maxiter is None else maxiter)
        prm.hypre_solver_i = int(P.hypre_solver_i)
        prm.mg_npre, prm.mg_npost, prm.mg_nbottom = mg_npre, mg_npost, mg_nbottom
        prm.mg_generic = int(mg_generic)
        prm.plane_solver = {"auto": 0, "mg": 1, "direct": 2}[solver]
        h = C.c_void_p()
        if self.lib.snac_create(C.byref(prm), C.byref(h)) != 0:
            raise L.SnacError(self.lib.snac_last_error(None).decode())
        self.h = h
        self._keep = []
        # every box of the run is described to the library; device state is created for the owned ones
        for r in range(W.nrank):
            self._add_box(r)
        self._chk(self.lib.snac_commit(self.h))
        if nproc > 1:
            from .dist import gather_blobs
            n = int(self.lib.snac_comm_blob_size())
            buf = C.create_string_buffer(n)
            self._chk(self.lib.snac_comm_export(self.h, buf))
            allb = gather_blobs(bytes(buf.raw), nproc)
            allbuf = C.create_string_buffer(allb, len(allb))
            self._chk(self.lib.snac_comm_connect(self.h, C.cast(allbuf, C.c_void_p), nproc))
        self.time, self.istep, self.dt, self.dt_cfl = 0.0, 0, 0.0, 0.0
        self.divtot = self.divmax = 0.0

    # -- plumbing --------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise L.SnacError(self.lib.snac_last_error(self.h).decode())

    def _add_box(self, r):
        P, W = self.case.params, self.case.world
        box, g, blk = W.boxes[r], self.case.grids[r], P.blocks[W.boxes[r].block]
        b = L.snac_box()
        b.rank, b.block = r, int(box.block)
        for d in range(3):
            b.lo[d], b.hi[d] = int(box.lo[d]), int(box.hi[d])
            for ib in range(2):
                b.nb[d][ib] = int(box.nb[d, ib])
                b.is_bound[d][ib] = int(box.is_bound[d, ib])
                b.cbcpre[d][ib] = blk.cbcpre[ib, d].encode()
                b.bcpre[d][ib] = float(blk.bcpre[ib, d])
                b.dl[d][ib] = float(g.dl[d, ib])
                bg = self.case.bgrids[box.block][d]
                b.dlf[d][ib] = float(bg.drf[1] if ib == 0 else bg.drf[-2])
                for c in range(3):
                    b.cbcvel[c][d][ib] = blk.cbcvel[ib, d, c].encode()
                    b.bcvel[c][d][ib] = float(blk.bcvel[ib, d, c])
                for s in range(P.nscal):
                    b.cbcscal[s][d][ib] = blk.cbcscal[ib, d, s].encode()
                    b.bcscal[s][d][ib] = float(blk.bcscal[ib, d, s])
        for name in ("dxc", "dxf", "dyc", "dyf", "dzc", "dzf"):
            a = np.ascontiguousarray(getattr(g, name), dtype=np.float64)
            self._keep.append(a)
            setattr(b, name, a.ctypes.data_as(C.POINTER(C.c_double)))
        self._chk(self.lib.snac_add_box(self.h, C.byref(b)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.snac_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shape(self, r):
        return tuple(int(x) + 2 for x in self.case.world.boxes[r].n)

    def upload(self, r, field, arr):
        a = np.asfortranarray(arr, dtype=np.float64)
        assert a.shape == self.shape(r), (a.shape, self.shape(r))
        self._chk(self.lib.snac_upload(self.h, r, field, a.ctypes.data_as(C.c_void_p)))

    def download(self, r, field):
        a = np.zeros(self.shape(r), order="F")
        self._chk(self.lib.snac_download(self.h, r, field, a.ctypes.data_as(C.c_void_p)))
        return a

    def download_async(self, r, field):
        """Starts the copy of what the field holds now into pinned host memory and returns ``(array, ticket, free)``: the
        time loop may go on; ``download_wait(ticket)`` (any thread) blocks until ``array`` is complete, ``free()`` releases
        the pinned memory."""
        shp = self.shape(r)
        n = int(np.prod(shp))
        ptr = self.lib.snac_host_alloc(n * 8)
        if not ptr:
            raise L.SnacError("snac_host_alloc failed")
        a = np.ctypeslib.as_array((C.c_double * n).from_address(ptr)).reshape(shp, order="F")
        t = C.c_int(-1)
        self._chk(self.lib.snac_download_async(self.h, r, field, C.c_void_p(ptr), C.byref(t)))
        lib = self.lib
        return a, t.value, (lambda: lib.snac_host_free(C.c_void_p(ptr)))

    def download_wait(self, ticket):
        if self.lib.snac_download_wait(self.h, ticket) != 0:
            raise L.SnacError("snac_download_wait failed")

    def set_fields(self, ini):
        """ini[r] = dict(u=,v=,w=,p=,s=[...]) of ghosted Fortran-order arrays."""
        for r in self.ranks:
            for key, fid in (("u", L.FIELD_U), ("v", L.FIELD_V), ("w", L.FIELD_W), ("p", L.FIELD_P)):
                self.upload(r, fid, ini[r][key])
            for i, s in enumerate(ini[r].get("s", [])):
                self.upload(r, L.FIELD_S0 + i, s)

    def get_fields(self, r):
        return dict(u=self.download(r, L.FIELD_U), v=self.download(r, L.FIELD_V), w=self.download(r, L.FIELD_W),
                    p=self.download(r, L.FIELD_P))

    # -- the reference's subroutines -----------------------------------------------------------
    def rk_mom(self, irk, dt=None):
        self._chk(self.lib.snac_rk_mom(self.h, irk, self.dt if dt is None else dt))

    def bounduvw(self, is_correc):
        self._chk(self.lib.snac_bounduvw(self.h, int(bool(is_correc))))

    def boundp(self, field=L.FIELD_P):
        self._chk(self.lib.snac_boundp(self.h, field))

    def set_inflow(self, planes):
        """planes[r][(idir, ib)] = (veln, velt1, velt2) as returned by ``case.inflow_planes``
        (u_in/v_in/w_in of the reference, src/main.f90:594-632)."""
        for r in self.ranks:
            for (idir, ib), vel in planes.get(r, {}).items():
                a = [np.asfortranarray(v, dtype=np.float64) for v in vel]
                self._chk(self.lib.snac_set_inflow(self.h, r, idir, ib, *[x.ctypes.data_as(C.c_void_p) for x in a]))

    def inflow(self, is_correc):
        self._chk(self.lib.snac_inflow(self.h, int(bool(is_correc))))

    def fillps(self, dtrk):
        self._chk(self.lib.snac_fillps(self.h, dtrk))

    def updt_rhs(self):
        self._chk(self.lib.snac_updt_rhs(self.h))

    def fft(self, direction):
        self._chk(self.lib.snac_fft(self.h, direction))

    def solve_n_helmholtz_2d(self):
        self._chk(self.lib.snac_solve_n_helmholtz_2d(self.h))

    def solve_helmholtz(self):
        self._chk(self.lib.snac_solve_helmholtz(self.h))

    def correc(self, dtrk):
        self._chk(self.lib.snac_correc(self.h, dtrk))

    def updt_pressure(self, alpha):
        self._chk(self.lib.snac_updt_pressure(self.h, alpha))

    def chkdt(self):
        v = C.c_double()
        self._chk(self.lib.snac_chkdt(self.h, C.byref(v)))
        return v.value

    def chkdiv(self):
        a, b = C.c_double(), C.c_double()
        self._chk(self.lib.snac_chkdiv(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def solver_stats(self):
        it, rr = C.c_int(), C.c_double()
        self._chk(self.lib.snac_solver_stats(self.h, C.byref(it), C.byref(rr)))
        return it.value, rr.value

    def solver_mean_iters(self):
        v = C.c_double()
        self._chk(self.lib.snac_solver_mean_iters(self.h, C.byref(v)))
        return v.value

    def solver_name(self):
        return {1: "mg", 2: "direct"}[int(self.lib.snac_plane_solver(self.h))]

    def launch_count(self):
        return int(self.lib.snac_launch_count(self.h))

    def sync(self):
        self._chk(self.lib.snac_sync(self.h))

    # -- driver logic (src/main.f90) -------------------------------------------------------------
    def init_bc(self):
        """src/main.f90:562-568,633-635."""
        self.bounduvw(False)
        self.boundp(L.FIELD_P)
        for i in range(self.case.params.nscal):
            self.boundp(L.FIELD_S0 + i)
        self.inflow(False)

    def set_dt(self):
        """src/main.f90:660-662,1051-1052."""
        P = self.case.params
        self.dt_cfl = self.chkdt()
        self.dt = P.dt_f if P.dt_f > 0.0 else min(P.cfl * self.dt_cfl, P.dtmax)
        return self.dt

    def advance(self, dt=None):
        """The `do irk=1,3` body on the device (src/main.f90:906-1033), nothing else."""
        self._chk(self.lib.snac_step(self.h, self.dt if dt is None else dt))

    def step(self):
        """One pass of the time loop (src/main.f90:897-1069), output excluded."""
        P = self.case.params
        self.istep += 1
        self.time += self.dt
        self.advance()
        if P.icheck > 0 and self.istep % P.icheck == 0:
            self.set_dt()
            self.divtot, self.divmax = self.chkdiv()
            if self.divtot != self.divtot:
                raise L.SnacError("ERROR: maximum divergence is too large.")   # src/main.f90:1063-1068
