"""Field/checkpoint files in the reference's on-disk layout, written from and read into device fields.

Twin of ``load_one``/``io_field`` (reference ``src/load.f90:84-187``), of ``out3d``/``write_log_output``
(``src/output.f90:227-303``) and of the ``geometry_b_###.out`` record (``src/main.f90:428-438``), so that
a run of this library and a run of the reference can restart from each other's ``fld_b_###_{u,v,w,p}.bin``
and the reference's post-processing readers (``utils/read_binary_data``) work unchanged.

One file per block and field: the block-global interior array ``ng(1) x ng(2) x ng(3)`` in Fortran
order, FP64, followed by two reals ``[time, istep]``.  The reference writes it with MPI-IO sub-array
views, one view per rank; here every box is written through a memory map at the same offsets, so
several processes (one per GPU) can write their boxes of the same file concurrently.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .initgrid import save_grid   # ``save_grid`` (src/initgrid.f90:128-143)


def _expected_size(ng) -> int:
    return (int(np.prod(np.asarray(ng, dtype=np.int64))) + 2) * 8


def _ensure_file(fname: str, size: int) -> None:
    """Creates the file if it is missing and gives it its size WITHOUT ever truncating existing contents: several
    processes (one per GPU) call this for the same file of a block and then write their own sub-arrays.  ``open('wb')``
    (O_TRUNC) here would wipe what a faster process has already written, or shrink the file under its live memory
    map; O_CREAT without O_TRUNC plus ftruncate to the one size every process agrees on is idempotent (the reference
    opens the file collectively and calls MPI_FILE_SET_SIZE, src/load.f90:141-143)."""
    fd = os.open(fname, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if os.fstat(fd).st_size != size:
            os.ftruncate(fd, size)
    finally:
        os.close(fd)


def out3d(fname: str, ng, lo, hi, p: np.ndarray, nh=(1, 1, 1)) -> None:
    """``out3d`` (src/output.f90:227-257): the block-global interior array, Fortran order, FP64, no trailer;
    each box writes its part (``io_field``, src/load.f90:161-187)."""
    ng = [int(x) for x in ng]
    n = [int(hi[d]) - int(lo[d]) + 1 for d in range(3)]
    ntot = int(np.prod(np.asarray(ng, dtype=np.int64)))
    _ensure_file(fname, ntot * 8)
    mm = np.memmap(fname, dtype=np.float64, mode="r+", shape=(ntot,))
    mm.reshape(ng, order="F")[tuple(slice(int(lo[d]) - 1, int(hi[d])) for d in range(3))] = \
        p[tuple(slice(nh[d], nh[d] + n[d]) for d in range(3))]
    mm.flush()
    del mm


def write_log_output(fname: str, fname_fld: str, varname: str, nmin, nmax, nskip, time: float, istep: int) -> None:
    """``write_log_output`` (src/output.f90:259-285): one line ``(A,A,A,9i5,E15.7,i10)`` per saved field, from
    which the reference's ``utils/visualize_fields`` scripts generate the xdmf file."""
    from .initgrid import _e15_7
    ints = "".join(f"{int(v):5d}" for v in (*nmin, *nmax, *nskip))
    with open(fname, "a") as f:
        f.write(f"{fname_fld} {varname}{ints}{_e15_7(float(time))}{int(istep):10d}\n")


def out3d_hook(sim, datadir: str, fldnum: str, ranks=None) -> None:
    """The calls of ``out3d.h90`` (``write_visu_3d`` for u,v,w,p and the scalars, src/output.f90:287-303):
    ``vex|vey|vez|pre_fld_<fldnum>_b_###.bin`` + a line each in ``log_visu_3d_b_###.out``."""
    from . import lib as L
    os.makedirs(datadir, exist_ok=True)
    names = [("vex", "Velocity_X", L.FIELD_U), ("vey", "Velocity_Y", L.FIELD_V), ("vez", "Velocity_Z", L.FIELD_W),
             ("pre", "Pressure_P", L.FIELD_P)]
    names += [(f"sca_{i + 1:03d}", f"Scalar_S{i + 1:03d}", L.FIELD_S0 + i) for i in range(sim.case.params.nscal)]
    for r in (sim.ranks if ranks is None else ranks):
        ib, ng, lo1, hi1, first = _block_frame(sim.case, r)
        cblock = f"{ib + 1:03d}"
        for stem, varname, fid in names:
            fname_bin = f"{stem}_fld_{fldnum}_b_{cblock}.bin"
            out3d(os.path.join(datadir, fname_bin), ng, lo1, hi1, sim.download(r, fid))
            if first:   # myid_block == 0
                write_log_output(os.path.join(datadir, f"log_visu_3d_b_{cblock}.out"), fname_bin, varname, (1, 1, 1),
                                 tuple(int(x) for x in ng), (1, 1, 1), sim.time, sim.istep)


def load_one(io: str, filename: str, ng, lo, hi, p: np.ndarray, nh=(1, 1, 1), time: float = 0.0, istep: int = 0,
             write_info: bool = True) -> Optional[Tuple[float, int]]:
    """``load_one`` (src/load.f90:84-160).  ``lo``/``hi``: block-local 1-based bounds of the box, ``p``: the
    box array with ``nh`` ghost layers.  'r' fills the interior of ``p`` and returns ``(time, istep)``;
    'w' writes the interior (and the trailer when ``write_info``: the reference lets rank 0 of the block do it)."""
    ng = [int(x) for x in ng]
    lo = [int(x) for x in lo]
    hi = [int(x) for x in hi]
    n = [hi[d] - lo[d] + 1 for d in range(3)]
    ntot = int(np.prod(np.asarray(ng, dtype=np.int64)))
    inner = tuple(slice(nh[d], nh[d] + n[d]) for d in range(3))
    box = tuple(slice(lo[d] - 1, hi[d]) for d in range(3))
    if io == "r":
        size = os.path.getsize(filename)
        if size != _expected_size(ng):   # src/load.f90:112-119
            raise IOError(f"checkpoint file with incorrect size: {filename} | expected size: {_expected_size(ng)} "
                          f"| actual size: {size}")
        mm = np.memmap(filename, dtype=np.float64, mode="r", shape=(ntot + 2,))
        p[inner] = mm[:ntot].reshape(ng, order="F")[box]
        info = (float(mm[ntot]), int(round(float(mm[ntot + 1]))))
        del mm
        return info
    if io == "w":
        _ensure_file(filename, _expected_size(ng))
        mm = np.memmap(filename, dtype=np.float64, mode="r+", shape=(ntot + 2,))
        mm[:ntot].reshape(ng, order="F")[box] = p[inner]
        if write_info:
            mm[ntot] = float(time)
            mm[ntot + 1] = float(istep)
        mm.flush()
        del mm
        return None
    raise ValueError("io must be 'r' or 'w'")


def save_geometry(datadir: str, case) -> None:
    """``geometry_b_###.out`` and ``grid_{x,y,z}_b_###.{bin,out}`` of every block (src/main.f90:428-438)."""
    os.makedirs(datadir, exist_ok=True)
    for ib, blk in enumerate(case.params.blocks):
        cblock = f"{ib + 1:03d}"
        boxes = [b for b in case.world.boxes if b.block == ib]
        lo_g = np.min([b.lo for b in boxes], axis=0)
        hi_g = np.max([b.hi for b in boxes], axis=0)
        for d, c in enumerate("xyz"):
            save_grid(os.path.join(datadir, f"grid_{c}_b_{cblock}"), case.bgrids[ib][d])
        with open(os.path.join(datadir, f"geometry_b_{cblock}.out"), "w") as f:
            f.write(" ".join(str(int(x)) for x in lo_g) + "\n")
            f.write(" ".join(str(int(x)) for x in hi_g) + "\n")
            f.write(" ".join(repr(float(x)) for x in blk.lmin) + "\n")
            f.write(" ".join(repr(float(x)) for x in blk.lmax) + "\n")


def _block_frame(case, r):
    """block index, block size ng and the block-local bounds lo_1/hi_1 of box r (src/main.f90:219-224)."""
    box = case.world.boxes[r]
    members = [b for b in case.world.boxes if b.block == box.block]
    lo_g = np.min([b.lo for b in members], axis=0)
    ng = case.params.blocks[box.block].ng
    first = min(b.rank for b in members) == box.rank   # myid_block == 0
    return box.block, ng, box.lo - lo_g + 1, box.hi - lo_g + 1, first


FIELD_SUFFIX = (("u", 0), ("v", 1), ("w", 2), ("p", 3))


def write_checkpoint(sim, datadir: str, basename: str = "fld", ranks=None) -> None:
    """The checkpoint of the time loop (src/main.f90:1090-1117): ``<basename>_b_###_{u,v,w,p}.bin`` (+ ``_s_###``)."""
    from . import lib as L
    os.makedirs(datadir, exist_ok=True)
    for r in (sim.ranks if ranks is None else ranks):
        ib, ng, lo1, hi1, first = _block_frame(sim.case, r)
        stem = os.path.join(datadir, f"{basename}_b_{ib + 1:03d}")
        for name, fid in FIELD_SUFFIX:
            load_one("w", f"{stem}_{name}.bin", ng, lo1, hi1, sim.download(r, fid), time=sim.time, istep=sim.istep,
                     write_info=first)
        for i in range(sim.case.params.nscal):
            load_one("w", f"{stem}_s_{i + 1:03d}.bin", ng, lo1, hi1, sim.download(r, L.FIELD_S0 + i), time=sim.time,
                     istep=sim.istep, write_info=first)


def write_checkpoint_async(sim, datadir: str, basename: str = "fld", ranks=None):
    """``write_checkpoint`` without stopping the time loop: the fields as they are NOW are packed on the device and copied to
    pinned host memory on a copy stream (``snac_download_async``); a writer thread waits for each copy and writes the same
    files.  Returns the thread: ``join()`` it before the files are read, linked or overwritten."""
    import threading

    from . import lib as L
    os.makedirs(datadir, exist_ok=True)
    time, istep = sim.time, sim.istep
    jobs = []
    for r in (sim.ranks if ranks is None else ranks):
        ib, ng, lo1, hi1, first = _block_frame(sim.case, r)
        stem = os.path.join(datadir, f"{basename}_b_{ib + 1:03d}")
        fields = [(f"{stem}_{name}.bin", fid) for name, fid in FIELD_SUFFIX]
        fields += [(f"{stem}_s_{i + 1:03d}.bin", L.FIELD_S0 + i) for i in range(sim.case.params.nscal)]
        for fname, fid in fields:
            arr, ticket, free = sim.download_async(r, fid)
            jobs.append((fname, ng, lo1, hi1, arr, ticket, free, first))

    def writer():
        for fname, ng, lo1, hi1, arr, ticket, free, first in jobs:
            sim.download_wait(ticket)
            load_one("w", fname, ng, lo1, hi1, arr, time=time, istep=istep, write_info=first)
            free()

    th = threading.Thread(target=writer, name="snac-checkpoint-writer")
    th.start()
    return th


def read_checkpoint(sim, datadir: str, basename: str = "fld", ranks=None) -> None:
    """Restart (src/main.f90:542-556): fields into the device arrays (ghost cells are filled by the first
    bounduvw/boundp, as in the reference), ``time`` and ``istep`` from the trailer."""
    from . import lib as L
    for r in (sim.ranks if ranks is None else ranks):
        ib, ng, lo1, hi1, _ = _block_frame(sim.case, r)
        stem = os.path.join(datadir, f"{basename}_b_{ib + 1:03d}")
        for name, fid in FIELD_SUFFIX:
            a = np.zeros(sim.shape(r), order="F")
            sim.time, sim.istep = load_one("r", f"{stem}_{name}.bin", ng, lo1, hi1, a)
            sim.upload(r, fid, a)
        for i in range(sim.case.params.nscal):
            a = np.zeros(sim.shape(r), order="F")
            load_one("r", f"{stem}_s_{i + 1:03d}.bin", ng, lo1, hi1, a)
            sim.upload(r, L.FIELD_S0 + i, a)
