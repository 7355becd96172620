"""Block -> rank decomposition and neighbour search.

Host-side twin of the reference's ``mod_initmpi`` (reference ``src/initmpi.f90:9-253``) and of the
per-rank block selection in ``read_input`` (``src/param.f90:207-228``). The reference computes this
with MPI all-gathers; every quantity is a deterministic function of ``blocks.nml``, so here every
process computes the description of *all* ranks (one rank = one box = one GPU) without
communication. Outputs per rank: ``lo, hi`` (global indices, 1-based like the reference),
``lo_g, hi_g`` (extents of the rank's block), ``nb[idir][ib]`` (neighbour rank or -1 where the
reference has ``MPI_PROC_NULL``), ``is_bound``; global: ``periods``.
"""
from __future__ import annotations

import dataclasses
from typing import List

import numpy as np

from .param import InputError, Params

PROC_NULL = -1


@dataclasses.dataclass
class Box:
    """One rank's box (reference: the per-rank state produced by ``initmpi``)."""
    rank: int
    block: int            # 0-based block index (my_block-1)
    id_first: int         # first rank of the block
    coords: np.ndarray    # rank coordinates inside the block, x fastest
    dims: np.ndarray
    ng: np.ndarray        # block size
    lo: np.ndarray        # global index of first interior cell (1-based)
    hi: np.ndarray
    lo_g: np.ndarray      # block extents in the global index space
    hi_g: np.ndarray
    nb: np.ndarray        # (3,2): nb[idir, ib]
    is_bound: np.ndarray  # (3,2) bool

    @property
    def n(self) -> np.ndarray:
        return self.hi - self.lo + 1


@dataclasses.dataclass
class World:
    boxes: List[Box]
    periods: np.ndarray      # (3,) int: global extent in periodic directions, else 0
    is_periodic: np.ndarray  # (3,) bool
    lo_g_min: np.ndarray
    hi_g_max: np.ndarray

    @property
    def nrank(self) -> int:
        return len(self.boxes)


def get_coords(idx: int, dims) -> np.ndarray:
    """x fastest (reference src/initmpi.f90:294-298)."""
    return np.array([idx % dims[0], (idx // dims[0]) % dims[1], (idx // (dims[0] * dims[1])) % dims[2]],
                    dtype=np.int64)


def get_id(coords, dims) -> int:
    if all(0 <= c <= d - 1 for c, d in zip(coords, dims)):
        return int(coords[0] + coords[1] * dims[0] + coords[2] * dims[0] * dims[1])
    return PROC_NULL


def block_origins(p: Params) -> np.ndarray:
    """Global index of each block's first cell, walking 'F' links from block 1
    (reference src/initmpi.f90:51-103). Returns lo_g[nblocks,3] (1-based)."""
    nb_ = p.nblocks
    lo_all = np.zeros((nb_, 3), dtype=np.int64)
    is_done = np.zeros(nb_, bool)
    is_unlocked = np.zeros(nb_, bool)
    iblock = 0
    lo_all[0, :] = 1
    is_unlocked[0] = True
    while not is_done.all():
        if iblock < 0:
            bad = int(np.argmax(~is_done)) + 1
            raise InputError(f"ERROR: invalid connectivity for block {bad}.")
        b = p.blocks[iblock]
        for idir in range(3):
            idir_t = [d for d in range(3) if d != idir]
            for inb in (0, 1):
                if b.cbcpre[inb, idir] == "F":
                    ifriend = int(round(b.bcpre[inb, idir])) - 1
                    if ifriend < 0 or ifriend >= nb_:
                        raise InputError(f"ERROR: invalid connectivity for block {iblock + 1}. "
                                         f"Block {ifriend + 1} does not exist.")
                    if not is_unlocked[ifriend]:
                        f = p.blocks[ifriend]
                        # periodicity check: a friend on the "wrong" side is a periodic wrap
                        if inb == 0 and f.lmin[idir] > b.lmin[idir]:
                            continue
                        if inb == 1 and f.lmin[idir] < b.lmin[idir]:
                            continue
                        ioffset = -(1 - inb) * int(f.ng[idir]) + inb * int(b.ng[idir])
                        lo_all[ifriend, idir] = lo_all[iblock, idir] + ioffset
                        lo_all[ifriend, idir_t] = lo_all[iblock, idir_t]
                        is_unlocked[ifriend] = True
        is_done[iblock] = True
        cand = np.nonzero(~is_done & is_unlocked)[0]
        iblock = int(cand[0]) if len(cand) else -1
    return lo_all


def initmpi(p: Params) -> World:
    """Decompose all blocks over ranks (reference ``initmpi``, src/initmpi.f90:9-253)."""
    lo_g_all = block_origins(p)
    hi_g_all = lo_g_all + np.array([b.ng for b in p.blocks]) - 1
    lo_g_min = lo_g_all.min(axis=0)
    hi_g_max = hi_g_all.max(axis=0)
    is_periodic = p.is_periodic
    periods = np.where(is_periodic, hi_g_max - lo_g_min + 1, 0).astype(np.int64)

    boxes: List[Box] = []
    id_first = 0
    for ib, b in enumerate(p.blocks):
        nr = int(np.prod(b.dims))
        for r in range(nr):
            coords = get_coords(r, b.dims)
            nb = np.full((3, 2), PROC_NULL, dtype=np.int64)
            for idir in range(3):
                e = np.zeros(3, dtype=np.int64)
                e[idir] = 1
                if coords[idir] - 1 >= 0:
                    nb[idir, 0] = id_first + get_id(coords - e, b.dims)
                if coords[idir] + 1 <= b.dims[idir] - 1:
                    nb[idir, 1] = id_first + get_id(coords + e, b.dims)
            n = b.ng // b.dims
            rem = b.ng % b.dims
            n = np.where(coords + 1 <= rem, n + 1, n)
            lo = lo_g_all[ib] + coords * n
            hi = lo_g_all[ib] - 1 + (coords + 1) * n
            shift = np.where(coords + 1 > rem, rem, 0)
            lo = lo + shift
            hi = hi + shift
            boxes.append(Box(rank=id_first + r, block=ib, id_first=id_first, coords=coords,
                             dims=b.dims.copy(), ng=b.ng.copy(), lo=lo.astype(np.int64),
                             hi=hi.astype(np.int64), lo_g=lo_g_all[ib].copy(), hi_g=hi_g_all[ib].copy(),
                             nb=nb, is_bound=np.zeros((3, 2), bool)))
        id_first += nr

    nrank = len(boxes)
    # cross-block / periodic neighbours (reference src/initmpi.f90:150-222)
    for me in boxes:
        b = p.blocks[me.block]
        for idir in range(3):
            for inb in (0, 1):
                if b.cbcpre[inb, idir] == "F":
                    fr = int(round(b.bcpre[inb, idir]))
                    if fr < 1 or fr > p.nblocks:
                        raise InputError(f"ERROR: invalid connectivity for block {me.block + 1}. "
                                         f"Block {fr} does not exist.")
                if me.nb[idir, inb] == PROC_NULL and b.cbcpre[inb, idir] == "F":
                    for other in boxes:
                        ob = p.blocks[other.block]
                        if not (ob.cbcpre[1 - inb, idir] == "F"
                                and (me.block + 1) == int(round(ob.bcpre[1 - inb, idir]))):
                            continue
                        is_nb = True
                        for iidir in range(3):
                            if iidir != idir:
                                is_nb = (is_nb and me.lo[iidir] == other.lo[iidir]
                                         and me.hi[iidir] == other.hi[iidir]
                                         and b.lmin[iidir] == ob.lmin[iidir]
                                         and b.lmax[iidir] == ob.lmax[iidir]
                                         and b.gr[iidir] == ob.gr[iidir]
                                         and b.gt[iidir] == ob.gt[iidir])
                        if not is_nb:
                            continue
                        if inb == 0 and me.lo[idir] == me.lo_g[idir]:
                            if me.lo[idir] == other.hi[idir] + 1:
                                me.nb[idir, inb] = other.rank
                            elif me.lo[idir] < other.hi[idir] + 1:
                                if periods[idir] == other.hi[idir] - me.lo[idir] + 1:
                                    me.nb[idir, inb] = other.rank
                                elif other.hi[idir] == other.hi_g[idir]:
                                    raise InputError(
                                        "ERROR: Inconsistent periodic boundary condition (?). "
                                        f"Expected: periods({idir + 1}) = {other.hi[idir] - me.lo[idir] + 1} "
                                        f"Found: {periods[idir]} (block {me.block + 1})")
                        elif inb == 1 and me.hi[idir] == me.hi_g[idir]:
                            if me.hi[idir] == other.lo[idir] - 1:
                                me.nb[idir, inb] = other.rank
                            elif me.hi[idir] > other.lo[idir] - 1:
                                if periods[idir] == me.hi[idir] - other.lo[idir] + 1:
                                    me.nb[idir, inb] = other.rank
                                elif other.lo[idir] == other.lo_g[idir]:
                                    raise InputError(
                                        "ERROR: Inconsistent periodic boundary condition (?). "
                                        f"Expected: periods({idir + 1}) = {me.hi[idir] - other.lo[idir] + 1} "
                                        f"Found: {periods[idir]} (block {me.block + 1})")
                        else:
                            raise InputError(f"ERROR: Expected connectivity between blocks {me.block + 1} "
                                             f"and {other.block + 1} not found.")
    for me in boxes:
        b = p.blocks[me.block]
        for idir in range(3):
            for inb in (0, 1):
                if b.cbcpre[inb, idir] == "F" and me.nb[idir, inb] == PROC_NULL:
                    raise InputError(f"ERROR: Expected connectivity between blocks {me.block + 1} and "
                                     f"{int(round(b.bcpre[inb, idir]))} is not possible. "
                                     "E.g. Blocks must share the same boundaries.")
        me.is_bound = me.nb == PROC_NULL
    assert nrank == p.nrank
    return World(boxes=boxes, periods=periods, is_periodic=is_periodic.copy(),
                 lo_g_min=lo_g_min, hi_g_max=hi_g_max)
