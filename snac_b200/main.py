"""``program snac`` with the time loop on the device: reads ``dns.nml`` / ``blocks.nml`` from a run directory and
writes ``data/`` in the reference's layout, so that an example case of the reference runs unchanged and its
post-processing (``utils/``, the readers used by ``tests/*/test.py``) works on the result.

Mirrors ``src/main.f90``: start-up and restart (``:195-253,428-438,542-568,633-662``), the time loop with its stop
criteria, ``chkdt``/``chkdiv`` every ``icheck`` steps and the output cadences (``:893-1144``).  All numerics are in
``libsnacb200.so`` (``snac_step``, ``snac_chkdt``, ``snac_chkdiv``, ``snac_out1d``); this file is control flow and files.

    python -m snac_b200.main RUNDIR [--fft-axis {0,1,2,3,auto}] [--impdiff] [--boussinesq]

The compile-time switches of the reference (``FFT_AXIS``, ``_IMPDIFF``, ``_BOUSSINESQ_BUOYANCY``) are command-line
options here.  One process drives one GPU; with ``torchrun`` every process owns the boxes ``rank % nproc``.
"""
from __future__ import annotations

import argparse
import os
import sys
import time as _time

from . import load, output
from .case import inflow_planes, initial_fields, setup_case
from .dist import allreduce, barrier
from .initgrid import _e15_7
from .param import SMALL   # epsilon*10**(precision/2), src/param.f90:9
from .sim import Simulation


def out0d(fname, var):
    """``out0d`` (src/output.f90:10-32): appends the values as ``(nE15.7)``."""
    with open(fname, "a") as f:
        f.write("".join(_e15_7(float(v)) for v in var) + "\n")


def run(rundir, fft_axis=0, impdiff=False, boussinesq=False, device=0, nproc=1, myproc=0, log=sys.stdout):
    case = setup_case(rundir, fft_axis=fft_axis, impdiff=impdiff, boussinesq=boussinesq)
    P = case.params
    datadir = os.path.join(rundir, "data")
    os.makedirs(datadir, exist_ok=True)
    owner = [r % nproc for r in range(case.world.nrank)]
    sim = Simulation(case, device=device, nproc=nproc, myproc=myproc, rank_owner=owner)
    root = myproc == 0
    if root:
        load.save_geometry(datadir, case)                       # src/main.f90:428-438
    if P.restart:                                               # src/main.f90:542-556
        load.read_checkpoint(sim, datadir)
        if root:
            print(f"*** Checkpoint loaded at time = {sim.time} time step = {sim.istep}. ***", file=log)
    else:
        sim.set_fields(initial_fields(case, ranks=sim.ranks))
    flags, planes = inflow_planes(case, ranks=sim.ranks)        # src/main.f90:594-632
    if any(f.any() for f in flags.values()):
        sim.set_inflow(planes)
    sim.init_bc()
    fldnum = f"{sim.istep:07d}"                                  # post-process and write the initial condition
    output.out1d_hook(sim, datadir, fldnum)                     # src/main.f90:654-656
    load.out3d_hook(sim, datadir, fldnum)
    sim.set_dt()                                                # src/main.f90:660-662
    if root:
        print(f"dt_cfl = {sim.dt_cfl} dt = {sim.dt}", file=log)
    savecounter, is_done, kill = 0, False, False
    writer = None   # the checkpoint being written by a thread while the time loop goes on (load.write_checkpoint_async)

    def finish_checkpoint():
        """joins the writer, then what needs every box of the files on disk: the barrier and the symbolic links"""
        nonlocal writer
        if writer is None:
            return
        th, base, stamp = writer
        writer = None
        th.join()
        barrier()                                               # every box of a file is on disk before anything reads or links it
        if not P.is_overwrite_save and root:                    # fld_* -> last checkpoint (symbolic links)
            for ib in range(len(P.blocks)):
                for suffix in ["u", "v", "w", "p"] + [f"s_{i + 1:03d}" for i in range(P.nscal)]:
                    link = os.path.join(datadir, f"fld_b_{ib + 1:03d}_{suffix}.bin")
                    if os.path.lexists(link):
                        os.remove(link)
                    os.symlink(f"{base}_b_{ib + 1:03d}_{suffix}.bin", link)
        if root:
            print(f"*** Checkpoint saved at time = {stamp[0]} time step = {stamp[1]}. ***", file=log)

    twi = _time.time()
    while not is_done:                                          # src/main.f90:893-1144
        sim.istep += 1
        sim.time += sim.dt
        if root:
            print(f"Timestep # {sim.istep} Time = {sim.time}", file=log)
        sim.advance()
        if P.stop_type[0] and sim.istep >= P.nstep:
            is_done = True
        if P.stop_type[1] and sim.time >= P.time_max:
            is_done = True
        if P.stop_type[2]:   # the slowest process decides (MPI_ALLREDUCE(MAX) of tw, src/main.f90:1043-1046):
            tw = float(allreduce([(_time.time() - twi) / 3600.0], "max")[0])   # all processes leave the loop together
            if tw >= P.tw_max:
                is_done = True
        if P.icheck > 0 and sim.istep % P.icheck == 0:
            sim.set_dt()
            if root:
                print(f"dt_cfl = {sim.dt_cfl} dt = {sim.dt}", file=log)
            if sim.dt_cfl < SMALL:
                print("ERROR: timestep is too small.\nAborting...", file=sys.stderr)
                is_done = kill = True
            sim.divtot, sim.divmax = sim.chkdiv()
            if root:
                print(f"Total divergence = {sim.divtot} | Maximum divergence = {sim.divmax}", file=log)
            if sim.divtot != sim.divtot:
                print("ERROR: maximum divergence is too large.\nAborting...", file=sys.stderr)
                is_done = kill = True
        fldnum = f"{sim.istep:07d}"
        if P.iout0d > 0 and sim.istep % P.iout0d == 0 and root:
            out0d(os.path.join(datadir, "time.out"), (1.0 * sim.istep, sim.dt, sim.time))
        if P.iout1d > 0 and sim.istep % P.iout1d == 0:
            output.out1d_hook(sim, datadir, fldnum)
        if P.iout2d > 0 and sim.istep % P.iout2d == 0:
            output.out2d_hook(sim, datadir, fldnum)
        if P.iout3d > 0 and sim.istep % P.iout3d == 0:
            load.out3d_hook(sim, datadir, fldnum)
        if P.isave > 0 and (sim.istep % P.isave == 0 or (is_done and not kill)):
            if P.is_overwrite_save:
                base = "fld"
            else:
                base = f"fld_{fldnum}"
                if P.nsaves_max > 0:
                    savecounter += 1
                    ichkpt = savecounter % P.nsaves_max + 1
                    base = f"fld_{ichkpt:04d}"
                    if root:
                        out0d(os.path.join(datadir, "log_saves.out"), (1.0 * sim.istep, sim.time, 1.0 * ichkpt))
            finish_checkpoint()                                 # (the previous one, if it is still being written)
            writer = (load.write_checkpoint_async(sim, datadir, basename=base), base, (sim.time, sim.istep))
            if is_done:
                finish_checkpoint()
    finish_checkpoint()
    return sim


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("rundir")
    ap.add_argument("--fft-axis", default="0", choices=("0", "1", "2", "3", "auto"),
                    help="FFT_AXIS of the reference's build; auto: a build the reference accepts for this input (transforms along a "
                         "uniform, undecomposed direction), else 0")
    ap.add_argument("--impdiff", action="store_true")
    ap.add_argument("--boussinesq", action="store_true")
    a = ap.parse_args(argv)
    nproc, myproc, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if nproc > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("gloo")
    sim = run(a.rundir, "auto" if a.fft_axis == "auto" else int(a.fft_axis), a.impdiff, a.boussinesq, device=local, nproc=nproc, myproc=myproc)
    sim.close()


if __name__ == "__main__":
    main()
