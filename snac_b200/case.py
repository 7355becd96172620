"""Case set-up shared by the driver: input files -> decomposition, grids, initial fields.

Mirrors the start-up part of ``program snac`` (reference ``src/main.f90:195-253,413-453,527-561``):
``read_input`` -> ``initmpi`` -> ``initgrid``/``load_grid`` -> ``distribute_grid``/``bound_grid`` ->
``initflow``/``initscal`` (or restart files).  Everything here is host-side, O(N) and executed once.
"""
from __future__ import annotations

import dataclasses
from typing import List

import numpy as np

from . import initflow as _if
from .initgrid import BoxGrid, Grid1D, block_grids, box_grids
from .initmpi import World, initmpi
from .param import Params, read_input


@dataclasses.dataclass
class Case:
    params: Params
    world: World
    bgrids: List[List[Grid1D]]
    grids: List[BoxGrid]
    rundir: str = "."


def setup_case(rundir: str = ".", fft_axis: int = 0, impdiff: bool = False, boussinesq: bool = False,
               params: Params = None) -> Case:
    auto = fft_axis in ("auto", -1)   # a build of the reference that accepts this input (param.pick_fft_axis)
    p = params if params is not None else read_input(rundir, 0 if auto else fft_axis, impdiff, boussinesq)
    if auto and params is None:
        from .param import pick_fft_axis
        p.fft_axis = pick_fft_axis(p)
    world = initmpi(p)
    bg = block_grids(p, rundir)
    g = box_grids(p, world, bg)
    return Case(params=p, world=world, bgrids=bg, grids=g, rundir=rundir)


def initial_fields(case: Case, ranks=None, noise: bool = False, mean_reduce=None):
    """Initial u,v,w,p (+scalars) for the given ranks (default: all), ghosted Fortran-order arrays.

    ``set_mean`` (reference src/initflow.f90:288-316) needs a sum over the boxes of a block
    (MPI_ALLREDUCE over comm_block there).  ``mean_reduce`` = None: the local means of *all* boxes of
    the block are evaluated here even when only some ranks are requested (one pass; fine for tests).
    Several processes at benchmark sizes pass ``mean_reduce``: a callable that sums an array of
    per-block partial means over the processes; each process then generates its own boxes only."""
    p, world = case.params, case.world
    ranks = list(range(world.nrank)) if ranks is None else list(ranks)
    out = {}
    need_blocks = {world.boxes[r].block for r in ranks}
    fields = {}
    for box in world.boxes:
        if box.block not in need_blocks or (mean_reduce is not None and box.rank not in ranks):
            continue
        b = p.blocks[box.block]
        l = b.lmax - b.lmin
        g = case.grids[box.rank]
        fields[box.rank] = _if.initflow(b.inivel, False, box, g, l, b.bcvel, p.visc, p.bforce[0],
                                        add_noise_to=noise)
    part = np.zeros(len(p.blocks))
    for r, f in fields.items():
        if f[4]:  # is_mean
            b = p.blocks[world.boxes[r].block]
            part[world.boxes[r].block] += _if.set_mean_local(case.grids[r].dxc, case.grids[r].dyf, case.grids[r].dzf,
                                                             b.lmax - b.lmin, f[0])
    if mean_reduce is not None:
        part = np.asarray(mean_reduce(part), dtype=np.float64)
    for r, f in fields.items():
        meanold = part[world.boxes[r].block]
        if f[4] and meanold != 0.0:
            f[0][...] = f[0] / meanold * f[5]
    for r in ranks:
        box = world.boxes[r]
        b = p.blocks[box.block]
        l = b.lmax - b.lmin
        u, v, w, pr = fields[r][:4]
        s = [_if.initscal(p.iniscal[i], b.bcscal[:, :, i], box, case.grids[r], l, bool(p.is_sforced[i]),
                          float(p.scalf[i])) for i in range(p.nscal)]
        out[r] = dict(u=u, v=v, w=w, p=pr, s=s)
    return out


def inflow_planes(case: Case, ranks=None):
    """Inflow planes of the boxes (reference src/main.f90:594-632): returns ``(is_bound_inflow, planes)``
    with ``is_bound_inflow[r]`` a (3,2) bool array ``[idir, ib]`` and ``planes[r][(idir, ib)] = (veln,
    velt1, velt2)``: ghosted planes of the normal and the two tangential components (x<y<z order)."""
    p, world = case.params, case.world
    ranks = list(range(world.nrank)) if ranks is None else list(ranks)
    flags, planes = {}, {}
    for r in ranks:
        box, g = world.boxes[r], case.grids[r]
        b = p.blocks[box.block]
        flags[r] = np.zeros((3, 2), bool)
        planes[r] = {}
        for idir in range(3):
            t = [d for d in range(3) if d != idir]
            for ib in range(2):
                on = bool(box.is_bound[idir, ib]) and int(b.inflow_type[ib, idir]) > 0 and b.cbcvel[ib, idir, idir] == "D"
                flags[r][idir, ib] = on
                if on:
                    planes[r][(idir, ib)] = _if.init_inflow(
                        int(b.inflow_type[ib, idir]), [world.periods[t[0]], world.periods[t[1]]],
                        [box.lo[t[0]], box.lo[t[1]]], [box.hi[t[0]], box.hi[t[1]]],
                        [b.lmin[t[0]], b.lmin[t[1]]], [b.lmax[t[0]], b.lmax[t[1]]],
                        g.rc(t[0]), g.rf(t[0]), g.rc(t[1]), g.rf(t[1]), float(b.bcvel[ib, idir, idir]), 1.0, p.visc)
    return flags, planes
