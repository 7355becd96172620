"""``out1d`` / ``out2d`` of the reference (src/output.f90:34-225) and the ``out1d.h90`` hook
(src/out1d.h90:17-52) with the averages evaluated on the device (``snac_out1d`` / ``snac_out2d``): only the
profile or the plane crosses PCIe.  Per block, as in the reference: the contributions of the boxes of a block
are added (``MPI_ALLREDUCE`` over ``comm_block``), the block's first box writes the text file."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import lib as L
from .dist import allreduce
from .initgrid import _e15_7


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _weights(sim, r, stag):
    """spacing arrays of box r for a variable staggered along ``stag`` (-1: cell-centred): the reference passes
    dxc for the staggered direction and dxf otherwise (src/out1d.h90)."""
    g = sim.case.grids[r]
    return [np.ascontiguousarray(g.dc(d) if d == stag else g.df(d), dtype=np.float64) for d in range(3)]


def block_profile(sim, block, field, idir, stag=-1):
    """p1d over the whole block (length ng[idir]), summed over the boxes of the block owned by this process."""
    P, W = sim.case.params, sim.case.world
    blk = P.blocks[block]
    l = np.asarray(blk.lmax, float) - np.asarray(blk.lmin, float)
    boxes = [b for b in W.boxes if b.block == block]
    lo_g = np.min([b.lo for b in boxes], axis=0)
    a, b_ = [d for d in range(3) if d != idir]
    out = np.zeros(int(blk.ng[idir]))
    for box in boxes:
        if box.rank not in sim.ranks:
            continue
        w = _weights(sim, box.rank, stag)
        p1d = np.zeros(int(box.n[idir]))
        sim._chk(sim.lib.snac_out1d(sim.h, box.rank, field, idir, _ptr(w[a]), _ptr(w[b_]), 1.0 / (l[a] * l[b_]), _ptr(p1d)))
        s = int(box.lo[idir] - lo_g[idir])
        out[s:s + p1d.size] += p1d
    return allreduce(out)   # MPI_ALLREDUCE over comm_block (src/output.f90:79-86): boxes of other processes


def block_average(sim, block, field, idir, stag=-1):
    """p2d over the whole block (shape ng[a] x ng[b]), the field averaged along idir."""
    P, W = sim.case.params, sim.case.world
    blk = P.blocks[block]
    l = np.asarray(blk.lmax, float) - np.asarray(blk.lmin, float)
    boxes = [b for b in W.boxes if b.block == block]
    lo_g = np.min([b.lo for b in boxes], axis=0)
    a, b_ = [d for d in range(3) if d != idir]
    out = np.zeros((int(blk.ng[a]), int(blk.ng[b_])), order="F")
    for box in boxes:
        if box.rank not in sim.ranks:
            continue
        w = _weights(sim, box.rank, stag)
        p2d = np.zeros((int(box.n[a]), int(box.n[b_])), order="F")
        sim._chk(sim.lib.snac_out2d(sim.h, box.rank, field, idir, _ptr(w[idir]), 1.0 / l[idir], _ptr(p2d)))
        sa, sb = int(box.lo[a] - lo_g[a]), int(box.lo[b_] - lo_g[b_])
        out[sa:sa + p2d.shape[0], sb:sb + p2d.shape[1]] += p2d
    return np.asfortranarray(allreduce(np.ascontiguousarray(out)))   # src/output.f90:170-177


def _is_block_root(sim, block):
    """myid_block == 0: the process owning the first box of the block writes its files (src/output.f90:87)."""
    first = min(b.rank for b in sim.case.world.boxes if b.block == block)
    return first in sim.ranks


def out1d(fname, sim, block, field, idir, stag=-1):
    """``out1d`` (src/output.f90:34-133): two columns ``(2E15.7)``: coordinate, profile.  Collective over the
    processes (every one calls it); the block's first process writes the file."""
    p1d = block_profile(sim, block, field, idir, stag)
    g = sim.case.bgrids[block][idir]
    coord = (g.rf if idir == stag else g.rc)[1:-1]
    if _is_block_root(sim, block):
        with open(fname, "w") as f:
            for x, v in zip(coord, p1d):
                f.write(_e15_7(float(x)) + _e15_7(float(v)) + "\n")
    return p1d


def out2d(fname, sim, block, field, idir, stag=-1):
    """``out2d`` (src/output.f90:135-225): three columns ``(3E15.7)``, the first kept direction fastest."""
    p2d = block_average(sim, block, field, idir, stag)
    a, b_ = [d for d in range(3) if d != idir]
    ga, gb = sim.case.bgrids[block][a], sim.case.bgrids[block][b_]
    ca = (ga.rf if a == stag else ga.rc)[1:-1]
    cb = (gb.rf if b_ == stag else gb.rc)[1:-1]
    if _is_block_root(sim, block):
        with open(fname, "w") as f:
            for jb in range(p2d.shape[1]):
                for ia in range(p2d.shape[0]):
                    f.write(_e15_7(float(ca[ia])) + _e15_7(float(cb[jb])) + _e15_7(float(p2d[ia, jb])) + "\n")
    return p2d


def out1d_hook(sim, datadir, fldnum):
    """The calls of ``out1d.h90`` (src/out1d.h90:17-52): profiles of u,v,w,p along x,y,z for every block."""
    os.makedirs(datadir, exist_ok=True)
    nblock = len(sim.case.params.blocks)
    for ib in range(nblock):
        cblock = f"{ib + 1:03d}"
        for name, field, stag in (("u", L.FIELD_U, 0), ("v", L.FIELD_V, 1), ("w", L.FIELD_W, 2), ("p", L.FIELD_P, -1)):
            for idir, c in enumerate("xyz"):
                out1d(os.path.join(datadir, f"{name}mean_{c}_fld_{fldnum}_b_{cblock}.out"), sim, ib, field, idir, stag)


def out2d_hook(sim, datadir, fldnum, enabled=False):
    """The calls of ``out2d.h90``.  The reference ships that include file with every call commented out
    (src/out2d.h90:14-25), so on the ``iout2d`` cadence it writes nothing; the same here unless ``enabled``, which
    writes the four z-averaged planes the commented calls name (``{u,v,w,p}mean_2d_z_fld_<fldnum>_b_###.out``)."""
    if not enabled:
        return
    os.makedirs(datadir, exist_ok=True)
    for ib in range(len(sim.case.params.blocks)):
        cblock = f"{ib + 1:03d}"
        for name, field, stag in (("u", L.FIELD_U, 0), ("v", L.FIELD_V, 1), ("w", L.FIELD_W, 2), ("p", L.FIELD_P, -1)):
            out2d(os.path.join(datadir, f"{name}mean_2d_z_fld_{fldnum}_b_{cblock}.out"), sim, ib, field, 2, stag)
