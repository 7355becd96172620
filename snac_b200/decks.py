"""Synthetic input decks (dns.nml / blocks.nml text) for the parity tests, ``smoke()`` and the
benchmark workloads, written in the reference's own namelist format (reference docs/INFO_INPUT.md)
and read back through ``snac_b200.param.read_input`` so that they exercise the same input path as a
real run.  The benchmark decks reproduce the reference's example cases at the sizes SURVEY.md
section 8(d) names: ``examples/_manuscript_taylor_green_vortex``, ``_manuscript_turbulent_channel``
and ``_manuscript_turbulent_duct``."""
from __future__ import annotations

import os

import numpy as np

from .case import setup_case


def _q(chars):
    return ", ".join(f"'{c}'" for c in chars)


def _f(vals):
    return ", ".join(repr(float(v)) for v in vals)


def write_case(rundir, blocks, visci=100.0, cfl=0.95, dt_f=-1.0, icheck=10, hypre_tol=None, hypre_maxiter=None,
               bforce=(0.0, 0.0, 0.0)):
    """blocks: list of dicts with keys dims, ng, lmin, lmax, gt, gr, cbcvel (3 strings of 6 chars:
    faces x-,x+,y-,y+,z-,z+ for u,v,w), cbcpre (6 chars), bcvel (3 lists of 6), bcpre (6), inivel."""
    os.makedirs(rundir, exist_ok=True)
    with open(os.path.join(rundir, "dns.nml"), "w") as f:
        f.write("&dns\n")
        f.write(f"cfl = {cfl}, dtmax = 1.e5, dt_f = {dt_f}\nvisci = {visci}\n")
        f.write("nstep = 10, time_max = 100., tw_max = 0.1\nstop_type(1:3) = T, F, F\n")
        f.write("restart = F, is_overwrite_save = T, nsaves_max = 0\n")
        f.write(f"icheck = {icheck}, iout0d = 10, iout1d = 0, iout2d = 0, iout3d = 0, isave = 0\n")
        f.write(f"bforce(1:3) = {_f(bforce)}\n/\n")
        if hypre_tol is not None or hypre_maxiter is not None:
            f.write(f"&hypre\nhypre_solver_i = 2, hypre_tol = {hypre_tol or 1e-4}, hypre_maxiter = {hypre_maxiter or 50}\n/\n")
    with open(os.path.join(rundir, "blocks.nml"), "w") as f:
        f.write(f"&blocks\nnblocks = {len(blocks)}\n")
        for ib, b in enumerate(blocks, start=1):
            f.write(f"block_dims(1:3,{ib}) = {', '.join(str(int(x)) for x in b['dims'])}\n")
            f.write(f"block_ng(1:3,{ib}) = {', '.join(str(int(x)) for x in b['ng'])}\n")
            f.write(f"block_lmin(1:3,{ib}) = {_f(b['lmin'])}\n")
            f.write(f"block_lmax(1:3,{ib}) = {_f(b['lmax'])}\n")
            f.write(f"block_gt(1:3,{ib}) = {', '.join(str(int(x)) for x in b.get('gt', (0, 0, 0)))}\n")
            f.write(f"block_gr(1:3,{ib}) = {_f(b.get('gr', (0, 0, 0)))}\n")
            for c in range(3):
                f.write(f"block_cbcvel(0:1,1:3,{c + 1},{ib}) = {_q(b['cbcvel'][c])}\n")
            f.write(f"block_cbcpre(0:1,1:3,{ib}) = {_q(b['cbcpre'])}\n")
            # on 'F' faces the boundary "value" is the 1-based id of the friend block
            # (reference docs/INFO_INPUT.md; src/initmpi.f90:76-98); default: the block itself
            friends = b.get("friends", [ib] * 6)
            bcv = [list(x) for x in b.get("bcvel", [[0.0] * 6] * 3)]
            bcp = list(b.get("bcpre", [0.0] * 6))
            for q in range(6):
                if b["cbcpre"][q] == "F":
                    bcp[q] = float(friends[q])
                    for c in range(3):
                        bcv[c][q] = float(friends[q])
            for c in range(3):
                f.write(f"block_bcvel(0:1,1:3,{c + 1},{ib}) = {_f(bcv[c])}\n")
            f.write(f"block_bcpre(0:1,1:3,{ib}) = {_f(bcp)}\n")
            f.write(f"block_inflow_type(0:1,1:3,{ib}) = {', '.join(str(int(x)) for x in b.get('inflow_type', [0] * 6))}\n")
            f.write(f"block_inivel({ib}) = '{b.get('inivel', 'zer')}'\n\n")
        f.write("/\n")


def channel_case(rundir, ng=(16, 12, 20), dims=(2, 1, 2), gr_z=2.0, **kw):
    """x,y periodic, no-slip walls in z (stretched): the turbulent-channel topology."""
    blk = dict(dims=dims, ng=ng, lmin=(0, 0, 0), lmax=(3.0, 1.5, 1.0), gt=(0, 0, 0), gr=(0, 0, gr_z),
               cbcvel=["FFFFDD"] * 3, cbcpre="FFFFNN", inivel="poi")
    write_case(rundir, [blk], **kw)


def cavity_case(rundir, ng=(12, 10, 14), dims=(1, 2, 1), **kw):
    """all walls, stretched in x and z, moving lid, one Dirichlet-pressure (outflow-like) face in y."""
    bcv = [[0.0] * 6 for _ in range(3)]
    bcv[0][5] = 1.0   # u = 1 on the z+ lid
    blk = dict(dims=dims, ng=ng, lmin=(0, 0, 0), lmax=(1.0, 1.2, 0.8), gt=(0, 0, 1), gr=(1.5, 0, 1.0),
               cbcvel=["DDDNDD", "DDDNDD", "DDDNDD"], cbcpre="NNNDNN", bcvel=bcv, inivel="zer")
    write_case(rundir, [blk], **kw)


def tgv_case(rundir, n=32, dims=(1, 1, 1), **kw):
    L = 2.0 * np.pi
    blk = dict(dims=dims, ng=(n, n, n), lmin=(0, 0, 0), lmax=(L, L, L), cbcvel=["FFFFFF"] * 3, cbcpre="FFFFFF",
               inivel="tgv")
    kw.setdefault("visci", 1600.0)
    write_case(rundir, [blk], **kw)


def duct_blocks_case(rundir, n=(8, 8, 8), **kw):
    """Four friend blocks 2x2 in y-z, x periodic, walls outside: the multi-block duct topology."""
    nx, ny, nz = n
    blocks = []
    for kz in range(2):
        for jy in range(2):
            pre = ["F", "F", "N" if jy == 0 else "F", "F" if jy == 0 else "N", "N" if kz == 0 else "F", "F" if kz == 0 else "N"]
            vel = ["F" if c == "F" else "D" for c in pre]
            me = 1 + jy + 2 * kz
            friends = [me, me, me - 1, me + 1, me - 2, me + 2]
            blocks.append(dict(dims=(1, 1, 1), ng=(nx, ny, nz), lmin=(0.0, 0.5 * jy, 0.5 * kz), friends=friends,
                               lmax=(1.0, 0.5 * (jy + 1), 0.5 * (kz + 1)), cbcvel=["".join(vel)] * 3,
                               cbcpre="".join(pre), inivel="poi"))
    write_case(rundir, blocks, **kw)


def lshape_blocks_case(rundir, n=(8, 8, 8), **kw):
    """Three friend blocks forming an L in the y-z cross-section (x periodic, walls outside): a multi-block
    geometry that is NOT a rectangle, so the planes of the per-wavenumber problems stay split over blocks and the
    solver has to couple them through the block interfaces (general multigrid path)."""
    nx, ny, nz = n
    #            y-   y+   z-   z+
    lay = {1: ("N", "F", "N", "F", (0.0, 0.0)),    # corner block: friends 2 (y+) and 3 (z+)
           2: ("F", "N", "N", "N", (0.5, 0.0)),    # right arm
           3: ("N", "N", "F", "N", (0.0, 0.5))}    # upper arm
    fr = {1: [1, 1, 1, 2, 1, 3], 2: [2, 2, 1, 2, 2, 2], 3: [3, 3, 3, 3, 1, 3]}
    blocks = []
    for ib in (1, 2, 3):
        ym, yp, zm, zp, (y0, z0) = lay[ib]
        pre = ["F", "F", ym, yp, zm, zp]
        vel = ["F" if c == "F" else "D" for c in pre]
        blocks.append(dict(dims=(1, 1, 1), ng=(nx, ny, nz), lmin=(0.0, y0, z0), lmax=(1.0, y0 + 0.5, z0 + 0.5), friends=fr[ib],
                           cbcvel=["".join(vel)] * 3, cbcpre="".join(pre), inivel="poi"))
    write_case(rundir, blocks, **kw)


def developing_channel_case(rundir, ng=(16, 8, 12), dims=(2, 1, 2), inflow=1, **kw):
    """The reference's examples/developing_channel topology (inflow at x-, outflow with p = 0 at x+, y periodic,
    walls in z) with an inflow plane of the given type on x- (src/main.f90:594-632; 1 = Poiseuille)."""
    bcv = [[0.0] * 6 for _ in range(3)]
    bcv[0][0] = 1.0   # bulk velocity of the inflow profile
    blk = dict(dims=dims, ng=ng, lmin=(0, 0, 0), lmax=(1.0, 1.5, 1.0), cbcvel=["DNFFDD", "NNFFDD", "NNFFDD"],
               cbcpre="NDFFNN", bcvel=bcv, inflow_type=[inflow, 0, 0, 0, 0, 0], inivel="zer")
    kw.setdefault("visci", 100.0)
    write_case(rundir, [blk], **kw)


def make(kind, rundir, fft_axis=0, impdiff=False, boussinesq=False, **kw):
    # the transform axis must not be decomposed (reference src/sanity.f90:75-79)
    if kind == "channel" and fft_axis == 1:
        kw.setdefault("dims", (1, 2, 2))
    if kind == "cavity" and fft_axis == 2:
        kw.setdefault("dims", (2, 1, 1))
    {"channel": channel_case, "cavity": cavity_case, "tgv": tgv_case, "duct": duct_blocks_case, "lshape": lshape_blocks_case,
     "devchannel": developing_channel_case}[kind](rundir, **kw)
    return setup_case(rundir, fft_axis=fft_axis, impdiff=impdiff, boussinesq=boussinesq)


def random_fields(case, seed=0, amp=1.0):
    """Random ghosted fields per rank (ghosts random too: the BC routines must overwrite them)."""
    rng = np.random.default_rng(seed)
    out = {}
    for r, box in enumerate(case.world.boxes):
        shp = tuple(int(x) + 2 for x in box.n)
        out[r] = {k: np.asfortranarray(amp * rng.standard_normal(shp)) for k in ("u", "v", "w", "p")}
    return out


# ---- benchmark workloads ---------------------------------
CHANNEL_DIMS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 1, 2), 8: (4, 1, 2)}   # the transform axis y stays whole


def bench_tgv(rundir, n=256, dims=(1, 1, 1), weak=False):
    """C2: examples/_manuscript_taylor_green_vortex on n^3 (per GPU when ``weak``), FFT_AXIS = 3."""
    L = 2.0 * np.pi
    mul = dims if weak else (1, 1, 1)
    blk = dict(dims=dims, ng=tuple(n * m for m in mul), lmin=(0, 0, 0), lmax=tuple(L * m for m in mul),
               cbcvel=["FFFFFF"] * 3, cbcpre="FFFFFF", inivel="tgv")
    write_case(rundir, [blk], visci=1600.0, cfl=0.95, icheck=10)
    return setup_case(rundir, fft_axis=3)


def bench_channel(rundir, ng=(1024, 512, 512), dims=(1, 1, 1), **kw):
    """C3: examples/_manuscript_turbulent_channel/{dns,blocks}.nml (x,y periodic, no-slip walls in z,
    lmax = 6,3,1, visci = 5640, Poiseuille start) on ``ng`` cells, FFT_AXIS = 2, one block split over
    ``dims`` boxes (never along y)."""
    blk = dict(dims=dims, ng=ng, lmin=(0, 0, 0), lmax=(6.0, 3.0, 1.0), cbcvel=["FFFFDD"] * 3, cbcpre="FFFFNN",
               inivel="poi")
    kw.setdefault("visci", 5640.0)
    write_case(rundir, [blk], cfl=0.95, icheck=10, **kw)
    return setup_case(rundir, fft_axis=2)


def bench_duct(rundir, ng=(512, 128, 128), nby=2, nbz=2, **kw):
    """C4: examples/_manuscript_turbulent_duct (x periodic, walls in y and z, visci = 4410) cut into
    nby x nbz friend blocks in the cross-section, FFT_AXIS = 1 (streamwise)."""
    nx, ny, nz = ng[0], ng[1] // nby, ng[2] // nbz
    blocks = []
    for kz in range(nbz):
        for jy in range(nby):
            pre = ["F", "F", "N" if jy == 0 else "F", "N" if jy == nby - 1 else "F", "N" if kz == 0 else "F",
                   "N" if kz == nbz - 1 else "F"]
            vel = ["F" if c == "F" else "D" for c in pre]
            me = 1 + jy + nby * kz
            friends = [me, me, me - 1, me + 1, me - nby, me + nby]
            blocks.append(dict(dims=(1, 1, 1), ng=(nx, ny, nz), lmin=(0.0, 1.0 * jy / nby, 1.0 * kz / nbz), friends=friends,
                               lmax=(10.0, 1.0 * (jy + 1) / nby, 1.0 * (kz + 1) / nbz), cbcvel=["".join(vel)] * 3,
                               cbcpre="".join(pre), inivel="poi"))
    kw.setdefault("visci", 4410.0)
    write_case(rundir, blocks, cfl=0.95, icheck=10, **kw)
    return setup_case(rundir, fft_axis=1)
