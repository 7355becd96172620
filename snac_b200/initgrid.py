"""1-D grid generation, distribution over boxes and inter-block ghost spacings.

Host-side twin of the reference's ``mod_initgrid`` (reference ``src/initgrid.f90:9-288``):
``initgrid`` (8 stretching maps, ``:171-287``; face/centre arrays incl. ghost rules ``:49-81``),
``distribute_grid`` (``:84-90``), ``bound_grid`` (``:91-127``; here a lookup into the neighbour's
arrays instead of MPI sends, the result is identical), ``save_grid``/``load_grid`` binary format
(``:128-165``). Arrays carry one ghost entry per side: python index ``q`` <-> Fortran ``lo-1+q``.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List

import numpy as np

from .initmpi import PROC_NULL, World
from .param import Params


def _two_end(alpha, r0):
    return 0.5 * (1.0 + math.tanh((r0 - 0.5) * alpha) / math.tanh(alpha / 2.0)) if alpha != 0.0 else r0


def _one_end(alpha, r0):
    return 1.0 * (1.0 + math.tanh((r0 - 1.0) * alpha) / math.tanh(alpha / 1.0)) if alpha != 0.0 else r0


def _one_end_r(alpha, r0):
    if alpha != 0.0:
        return 1.0 - 1.0 * (1.0 + math.tanh((1.0 - r0 - 1.0) * alpha) / math.tanh(alpha / 1.0))
    return r0


def _middle(alpha, r0):
    if alpha != 0.0:
        if r0 <= 0.5:
            return 0.5 * (1.0 - 1.0 + math.tanh(2.0 * alpha * (r0 - 0.0)) / math.tanh(alpha))
        return 0.5 * (1.0 + 1.0 + math.tanh(2.0 * alpha * (r0 - 1.0)) / math.tanh(alpha))
    return r0


def _geo_one_end(alpha, r0):
    power = 1.0 / 3.0
    if r0 == 1.0:
        return r0
    if r0 == 0.0:
        return 0.0
    return r0 * ((1.0 - r0 ** alpha) / (1.0 - r0) / alpha) ** power


def _geo_one_end_r(alpha, r0):
    return 1.0 - _geo_one_end(alpha, 1.0 - r0)


def _geo_two_ends(alpha, r0):
    if r0 <= 0.5:
        return _geo_one_end(alpha / 2.0, 2.0 * r0) / 2.0
    return 1.0 - _geo_one_end(alpha / 2.0, 2.0 * (1.0 - r0)) / 2.0


def _geo_middle(alpha, r0):
    if r0 <= 0.5:
        return (1.0 - _geo_one_end(alpha / 2.0, (0.5 - r0) * 2.0)) / 2.0
    return (1.0 + _geo_one_end(alpha / 2.0, (r0 - 0.5) * 2.0)) / 2.0


_MAPS = {0: _two_end, 1: _one_end, 2: _middle, 3: _one_end_r,
         4: _geo_one_end, 5: _geo_one_end_r, 6: _geo_two_ends, 7: _geo_middle}


@dataclasses.dataclass
class Grid1D:
    """drc (centre spacing), drf (face spacing = cell width), rc, rf with one ghost per side."""
    drc: np.ndarray
    drf: np.ndarray
    rc: np.ndarray
    rf: np.ndarray


def initgrid(n: int, gt: int, gr: float, lmin: float, lmax: float) -> Grid1D:
    """Reference ``initgrid`` (src/initgrid.f90:9-83) for a block of ``n`` cells."""
    fmap = _MAPS.get(int(gt), _two_end)
    rf = np.zeros(n + 2)
    drf = np.zeros(n + 2)
    drc = np.zeros(n + 2)
    rc = np.zeros(n + 2)
    for q in range(1, n + 1):
        r0 = (q - 0.0) / (1.0 * n)
        rf[q] = lmin + fmap(gr, r0) * (lmax - lmin)
    rf[0] = lmin
    for q in range(1, n + 1):
        drf[q] = rf[q] - rf[q - 1]
    drf[0] = drf[1]
    drf[n + 1] = drf[n]
    for q in range(0, n + 1):
        drc[q] = 0.5 * (drf[q] + drf[q + 1])
    drc[n + 1] = drc[n]
    rc[0] = lmin - drc[0] / 2.0
    rf[0] = lmin
    for q in range(1, n + 2):
        rc[q] = rc[q - 1] + drc[q - 1]
        rf[q] = rf[q - 1] + drf[q]
    return Grid1D(drc=drc, drf=drf, rc=rc, rf=rf)


def save_grid(fname: str, g: Grid1D) -> None:
    """``fname.bin``: one record rf,rc,drf,drc (interior); ``fname.out``: 5-column text
    (reference src/initgrid.f90:128-143)."""
    n = len(g.rf) - 2
    with open(fname + ".bin", "wb") as f:
        for a in (g.rf, g.rc, g.drf, g.drc):
            np.ascontiguousarray(a[1:n + 1], dtype=np.float64).tofile(f)
    with open(fname + ".out", "w") as f:
        for q in range(n + 2):
            f.write("".join(_e15_7(v) for v in (0.0, g.rf[q], g.rc[q], g.drf[q], g.drc[q])) + "\n")


def load_grid(fname: str, n: int) -> Grid1D:
    """Reference ``load_grid`` incl. its ghost reconstruction (src/initgrid.f90:144-165)."""
    a = np.fromfile(fname + ".bin", dtype=np.float64)
    if a.size != 4 * n:
        raise RuntimeError(f"grid file {fname}.bin has {a.size} reals, expected {4 * n}")
    rf, rc, drf, drc = (np.zeros(n + 2) for _ in range(4))
    rf[1:n + 1], rc[1:n + 1], drf[1:n + 1], drc[1:n + 1] = a[0:n], a[n:2 * n], a[2 * n:3 * n], a[3 * n:4 * n]
    drf[0] = drf[1]
    drf[n + 1] = drf[n]
    drc[0] = drf[1]
    drc[n + 1] = drc[n]
    rf[0] = rf[1] - drf[1]
    rf[n + 1] = rf[n] + drf[n]
    rc[0] = rc[1] - drc[0]
    rc[n + 1] = rc[n] + drc[n]
    return Grid1D(drc=drc, drf=drf, rc=rc, rf=rf)


def _e15_7(v: float) -> str:
    """Fortran ``E15.7`` edit descriptor (0.dddddddE+xx)."""
    if v == 0.0 or not math.isfinite(v):
        s = "0.0000000E+00" if v == 0.0 else str(v)
        return s.rjust(15)
    exp = int(math.floor(math.log10(abs(v)))) + 1
    man = abs(v) / 10.0 ** exp
    man = round(man, 7)
    if man >= 1.0:
        man /= 10.0
        exp += 1
    s = f"{'-' if v < 0 else ''}0.{int(round(man * 1e7)):07d}E{'+' if exp >= 0 else '-'}{abs(exp):02d}"
    return s.rjust(15)


@dataclasses.dataclass
class BoxGrid:
    """Per-box 1-D metric arrays, length n+2 each (index lo-1..hi+1), after ``bound_grid``."""
    dxc: np.ndarray
    dxf: np.ndarray
    xc: np.ndarray
    xf: np.ndarray
    dyc: np.ndarray
    dyf: np.ndarray
    yc: np.ndarray
    yf: np.ndarray
    dzc: np.ndarray
    dzf: np.ndarray
    zc: np.ndarray
    zf: np.ndarray
    dl: np.ndarray  # (3,2): d?c_g(lo_g-1), d?c_g(hi_g) of the box's block (reference src/main.f90:726-728)

    def dc(self, d):
        return (self.dxc, self.dyc, self.dzc)[d]

    def df(self, d):
        return (self.dxf, self.dyf, self.dzf)[d]

    def rc(self, d):
        return (self.xc, self.yc, self.zc)[d]

    def rf(self, d):
        return (self.xf, self.yf, self.zf)[d]


def block_grids(p: Params, rundir: str = ".") -> List[List[Grid1D]]:
    """Block-global grids [block][dir]; ``grid/grid_{x,y,z}_b_###.bin`` overrides the analytic map
    when present (reference src/main.f90:416-427)."""
    out = []
    for ib, b in enumerate(p.blocks):
        gs = []
        for d, c in enumerate("xyz"):
            g = initgrid(int(b.ng[d]), int(b.gt[d]), float(b.gr[d]), float(b.lmin[d]), float(b.lmax[d]))
            fn = os.path.join(rundir, "grid", f"grid_{c}_b_{ib + 1:03d}")
            if os.path.exists(fn + ".bin"):
                g = load_grid(fn, int(b.ng[d]))
            gs.append(g)
        out.append(gs)
    return out


def box_grids(p: Params, world: World, bgrids: List[List[Grid1D]]) -> List[BoxGrid]:
    """``distribute_grid`` + ``bound_grid`` for every box (reference src/main.f90:439-453,
    src/initgrid.f90:84-127)."""
    res: List[BoxGrid] = []
    parts: List[Dict[str, List[np.ndarray]]] = []
    for box in world.boxes:
        arrs = {"dc": [], "df": [], "rc": [], "rf": []}
        for d in range(3):
            g = bgrids[box.block][d]
            s = int(box.lo[d] - box.lo_g[d])          # python index of (lo-1) in the block-global array
            e = s + int(box.n[d]) + 2
            arrs["dc"].append(g.drc[s:e].copy())
            arrs["df"].append(g.drf[s:e].copy())
            arrs["rc"].append(g.rc[s:e].copy())
            arrs["rf"].append(g.rf[s:e].copy())
        parts.append(arrs)
    # bound_grid: ghost cell widths across block faces come from the neighbour's first/last cell;
    # values sent are the pre-exchange interior widths, so read them from a snapshot
    snap = [[a.copy() for a in parts[r]["df"]] for r in range(world.nrank)]
    for box, arrs in zip(world.boxes, parts):
        for d in range(3):
            n = int(box.n[d])
            f, c = arrs["df"][d], arrs["dc"][d]
            if box.lo[d] == box.lo_g[d]:
                nbr = int(box.nb[d, 0])
                if nbr != PROC_NULL:
                    nn = int(world.boxes[nbr].n[d])
                    f[0] = snap[nbr][d][nn]          # neighbour's grid_f(hi)
                c[0] = (f[1] + f[0]) / 2.0
            if box.hi[d] == box.hi_g[d]:
                nbr = int(box.nb[d, 1])
                if nbr != PROC_NULL:
                    f[n + 1] = snap[nbr][d][1]       # neighbour's grid_f(lo)
                c[n] = (f[n + 1] + f[n]) / 2.0
                c[n + 1] = c[n]
    for box, arrs in zip(world.boxes, parts):
        g = bgrids[box.block]
        dl = np.array([[g[d].drc[0], g[d].drc[int(box.ng[d])]] for d in range(3)])
        res.append(BoxGrid(dxc=arrs["dc"][0], dxf=arrs["df"][0], xc=arrs["rc"][0], xf=arrs["rf"][0],
                           dyc=arrs["dc"][1], dyf=arrs["df"][1], yc=arrs["rc"][1], yf=arrs["rf"][1],
                           dzc=arrs["dc"][2], dzf=arrs["df"][2], zc=arrs["rc"][2], zf=arrs["rf"][2],
                           dl=dl))
    return res
