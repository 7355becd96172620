"""Initial conditions for velocity, pressure and scalars.

Host-side twin of the reference's ``mod_initflow`` (reference ``src/initflow.f90:10-187`` velocity,
``:188-262`` scalars, ``:369-429`` inflow planes). Arrays are Fortran-ordered numpy arrays with one
ghost layer: python index ``[i,j,k]`` <-> Fortran ``(lo-1+i, lo-1+j, lo-1+k)``.

Noise: the reference draws from the compiler's ``random_number`` with seeds 123/456/789 walking the
whole block in index order (``src/initflow.f90:262-286``) -- not reproducible without that
compiler. Here the draw is a counter-based uniform (splitmix64 of seed and global linear index),
which keeps the reference's defining property: decomposition-independent values.
"""
from __future__ import annotations

import math

import numpy as np

from .param import PI, InputError


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_noise(seed: int, lo, hi, lo_g, hi_g) -> np.ndarray:
    """Uniform [0,1) per interior cell of the box, a function of (seed, global index) only."""
    ng = np.asarray(hi_g) - np.asarray(lo_g) + 1
    i = np.arange(lo[0], hi[0] + 1, dtype=np.uint64) - np.uint64(lo_g[0])
    j = np.arange(lo[1], hi[1] + 1, dtype=np.uint64) - np.uint64(lo_g[1])
    k = np.arange(lo[2], hi[2] + 1, dtype=np.uint64) - np.uint64(lo_g[2])
    lin = (i[:, None, None] + np.uint64(ng[0]) * (j[None, :, None] + np.uint64(ng[1]) * k[None, None, :]))
    with np.errstate(over="ignore"):
        h = _splitmix64(lin + (np.uint64(seed) << np.uint64(40)))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def add_noise(lo, hi, lo_g, hi_g, iseed: int, norm: float, p: np.ndarray) -> None:
    """Reference ``add_noise`` (src/initflow.f90:262-286) with the counter-based draw."""
    rn = uniform_noise(iseed, lo, hi, lo_g, hi_g)
    p[1:-1, 1:-1, 1:-1] += 2.0 * (rn - 0.5) * norm


def _poiseuille(zc, l, norm):
    z = zc / l
    return 6.0 * z * (1.0 - z) * norm


def _couette(zc, l, norm):
    z = zc / l
    return 0.5 * (1.0 - 2.0 * z) * norm


def _log_profile(zc, l, uref, lref, visc):
    reb = lref * uref / visc
    retau = 0.09 * reb ** 0.88
    z = zc * 2.0 * retau
    with np.errstate(divide="ignore", invalid="ignore"):
        pr = 2.5 * np.log(z) + 5.5
    return np.where(z <= 11.6, z, pr)


def set_mean_local(dx, dy, dz, l, p) -> float:
    """Local part of the mean in ``set_mean`` (src/initflow.f90:288-316); the caller sums it over the
    boxes of the block (the reference's allreduce on comm_block)."""
    w = dx[1:-1, None, None] * dy[None, 1:-1, None] * dz[None, None, 1:-1]
    return float(np.sum(p[1:-1, 1:-1, 1:-1] * w) / (l[0] * l[1] * l[2]))


def initflow(inivel: str, is_wallturb: bool, box, grid, l, bcvel, visc, bforce1, add_noise_to=None):
    """Returns (u,v,w,p, is_mean, ubulk) for one box; the caller applies ``set_mean`` over the block.

    Reference ``initflow`` (src/initflow.f90:10-187)."""
    n = box.n
    shp = (n[0] + 2, n[1] + 2, n[2] + 2)
    u = np.zeros(shp, order="F")
    v = np.zeros(shp, order="F")
    w = np.zeros(shp, order="F")
    p = np.zeros(shp, order="F")
    zc = grid.zc[1:-1]
    is_noise = is_mean = False
    uref = ubulk = lref = 1.0
    u1d = None
    key = inivel.strip()[:3]
    if key == "cou":
        u1d = _couette(zc, l[2], 1.0) + 0.5
        u1d = bcvel[0, 2, 0] * u1d + bcvel[1, 2, 0] * (1.0 - u1d)
    elif key == "poi":
        u1d = _poiseuille(zc, l[2], ubulk)
        is_mean = True
    elif key == "zer":
        u1d = np.zeros_like(zc)
    elif key == "uni":
        u1d = np.full_like(zc, uref)
    elif key == "log":
        u1d = _log_profile(zc, l[2], ubulk, l[2], visc)
        is_noise = is_mean = True
    elif key == "hcp":
        u1d = _poiseuille(zc, 2.0 * l[2], ubulk)
        is_mean = True
    elif key == "hcl":
        u1d = _log_profile(zc, 2.0 * l[2], ubulk, 2.0 * l[2], visc)
        is_noise = is_mean = True
    elif key == "tgv":
        xcl = grid.xc[1:-1] / l[0] * 2.0 * PI
        xfl = grid.xf[1:-1] / l[0] * 2.0 * PI
        ycl = grid.yc[1:-1] / l[1] * 2.0 * PI
        yfl = grid.yf[1:-1] / l[1] * 2.0 * PI
        zcl = grid.zc[1:-1] / l[2] * 2.0 * PI
        u[1:-1, 1:-1, 1:-1] = (np.sin(xfl)[:, None, None] * np.cos(ycl)[None, :, None]) * np.cos(zcl)[None, None, :]
        v[1:-1, 1:-1, 1:-1] = (-np.cos(xcl)[:, None, None] * np.sin(yfl)[None, :, None]) * np.cos(zcl)[None, None, :]
    elif key == "kov":
        reb = uref * lref / visc
        lam = reb / 2 - math.sqrt(reb ** 2 / 4 + 4 * PI ** 2)
        xc, xf = grid.xc[1:-1], grid.xf[1:-1]
        zcc, zf = grid.zc[1:-1], grid.zf[1:-1]
        u[1:-1, 1:-1, 1:-1] = (1.0 - np.exp(lam * xf)[:, None, None] * np.cos(2 * PI * zcc)[None, None, :]) \
            * np.ones(n[1])[None, :, None]
        w[1:-1, 1:-1, 1:-1] = (lam / (2.0 * PI) * np.exp(lam * xc)[:, None, None]
                               * np.sin(2 * PI * zf)[None, None, :]) * np.ones(n[1])[None, :, None]
    elif key == "pdc":
        lref = l[2] / 2.0
        if is_wallturb:
            uref = math.sqrt(bforce1 * lref)
            retau = uref * lref / visc
            reb = (retau / 0.09) ** (1.0 / 0.88)
            ubulk = reb * visc / (2.0 * lref)
        else:
            ubulk = bforce1 * lref ** 2 / (3.0 * visc)
        u1d = _poiseuille(zc, l[2], ubulk)
        is_mean = True
    else:
        raise InputError("ERROR: invalid name for initial velocity field")
    if key not in ("tgv", "kov"):
        u[1:-1, 1:-1, 1:-1] = u1d[None, None, :]
    if is_noise or add_noise_to:
        add_noise(box.lo, box.hi, box.lo_g, box.hi_g, 123, 0.05, u)
        add_noise(box.lo, box.hi, box.lo_g, box.hi_g, 456, 0.05, v)
        add_noise(box.lo, box.hi, box.lo_g, box.hi_g, 789, 0.05, w)
    return u, v, w, p, is_mean, ubulk


def initscal(iniscal: str, bcscal, box, grid, l, is_sforced: bool, scalf: float):
    """Reference ``initscal`` (src/initflow.f90:188-261). Returns the ghosted scalar array."""
    n = box.n
    s = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    zc = grid.zc[1:-1]
    sref = 0.5 * (bcscal[0, 2] + bcscal[1, 2])
    if is_sforced:
        sref = scalf
    key = iniscal.strip()[:3]
    if key in ("zer", "dhc"):
        s1d = np.zeros_like(zc)
    elif key == "uni":
        s1d = np.full_like(zc, sref)
    elif key == "cou":
        s1d = _couette(zc, l[2], 1.0) + 0.5
        s1d = bcscal[0, 2] * s1d + bcscal[1, 2] * (1.0 - s1d)
    else:
        raise InputError("ERROR: invalid name for initial scalar field")
    s[1:-1, 1:-1, 1:-1] = s1d[None, None, :]
    if key == "dhc":
        xx = grid.xc[1:-1] / l[0]
        s[1:-1, 1:-1, 1:-1] = ((1.0 - xx) * bcscal[0, 0] + xx * bcscal[1, 0])[:, None, None]
    return s


def init_inflow(inflow_type: int, periods2, lo2, hi2, lmin2, lmax2, x1c, x1f, x2c, x2f, uref, lref, visc):
    """Inflow planes (veln, velt1, velt2), ghosted in both plane directions
    (reference src/initflow.f90:369-429)."""
    n1, n2 = len(x1c), len(x2c)
    veln = np.zeros((n1, n2), order="F")
    velt1 = np.zeros((n1, n2), order="F")
    velt2 = np.zeros((n1, n2), order="F")

    def psq(r, a, b):
        rr = (r - a) / (b - a)
        return 6.0 * (rr * (1.0 - rr))

    if inflow_type == 1:
        q = [0 if periods2[0] != 0 else 1, 0 if periods2[1] != 0 else 1]
        f1 = psq(x1c, lmin2[0], lmax2[0]) ** q[0]
        f2 = psq(x2c, lmin2[1], lmax2[1]) ** q[1]
        veln[:, :] = uref * f1[:, None] * f2[None, :]
    elif inflow_type == 2:
        reb = uref * lref / visc
        lam = reb / 2 - math.sqrt(reb ** 2 / 4 + 4 * PI ** 2)
        x0 = -0.5
        veln[:, :] = (1.0 - math.exp(lam * x0) * np.cos(2 * PI * x2c))[None, :]
        velt2[:, :] = (lam / (2.0 * PI) * math.exp(lam * x0) * np.sin(2 * PI * x2f))[None, :]
    return veln, velt1, velt2
