"""Run-time parameters: `dns.nml` / `blocks.nml` namelist reader.

Host-side twin of the reference's ``mod_param`` (reference ``src/param.f90:68-277``): same
namelist groups (``&dns &scalar &hypre &blocks``), same defaults (``src/param.f90:110-132``,
``:152-158``), same derived quantities (``visc``, ``alpha_max``, ``is_periodic``, ``vol_all``).
The physics variants that the reference selects at compile time (``FFT_AXIS``, ``IMPDIFF``,
``BOUSSINESQ_BUOYANCY``; reference ``configs/defaults/build-default.conf:12-19``) are run-time
fields of :class:`Params` here.

No Fortran runtime exists in this image, so the namelist syntax actually used by the reference
inputs is parsed by hand: several ``name = values`` per line, array sections with explicit bounds
filled in column-major order, ``(:)`` sections, Fortran reals/logicals, quoted strings, ``!``
comments and ``r*value`` repeats.
"""
from __future__ import annotations

import dataclasses
import itertools
import math
import os
import re
from typing import Dict, List, Sequence, Tuple

import numpy as np

# --- constants (reference src/param.f90:8-19) ------------------------------------------------
PI = math.acos(-1.0)
SMALL = float(np.finfo(np.float64).eps * 10.0 ** (15 / 2.0))  # epsilon*10**(precision/2)
DATADIR = "data/"
RKCOEFF = np.array([[32.0 / 60.0, 0.0],
                    [25.0 / 60.0, -17.0 / 60.0],
                    [45.0 / 60.0, -25.0 / 60.0]])  # rkcoeff(:,irk) -> RKCOEFF[irk-1]
HYPRE_SOLVER_I_DEFAULT = 2
HYPRE_TOL_DEFAULT = float(np.float32(1.0e-4))  # real(1.e-4,rp): single-precision literal promoted
HYPRE_MAXITER_DEFAULT = 50
MAX_BLOCKS = 999


class InputError(RuntimeError):
    """Raised where the reference prints to stderr, finalizes MPI and ``error stop``s."""


# --- namelist tokenizer ---------------------------------------------------------------------
_ASSIGN_RE = re.compile(r"([A-Za-z_][A-Za-z0-9_]*)\s*(\([^)]*\))?\s*=")


def _strip_comments(text: str) -> str:
    out = []
    for line in text.splitlines():
        buf, q = [], None
        for ch in line:
            if q:
                buf.append(ch)
                if ch == q:
                    q = None
            elif ch in "'\"":
                q = ch
                buf.append(ch)
            elif ch == "!":
                break
            else:
                buf.append(ch)
        out.append("".join(buf))
    return "\n".join(out)


def _find_group(text: str, group: str):
    """Return the body of ``&group ... /`` or None (a missing group is legal for &hypre/&scalar)."""
    m = re.search(r"&" + group + r"\b", text, flags=re.IGNORECASE)
    if not m:
        return None
    i, q = m.end(), None
    while i < len(text):
        ch = text[i]
        if q:
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
        elif ch == "/":
            return text[m.end():i]
        i += 1
    raise InputError(f"namelist group &{group} is not terminated by '/'")


def _split_values(s: str) -> List[str]:
    vals, buf, q = [], [], None
    for ch in s:
        if q:
            buf.append(ch)
            if ch == q:
                q = None
        elif ch in "'\"":
            q = ch
            buf.append(ch)
        elif ch in ", \t\n\r":
            if buf:
                vals.append("".join(buf))
                buf = []
        else:
            buf.append(ch)
    if buf:
        vals.append("".join(buf))
    out = []
    for v in vals:  # r*c repeat counts
        m = re.fullmatch(r"(\d+)\*(.+)", v)
        if m and not v.startswith(("'", '"')):
            out.extend([m.group(2)] * int(m.group(1)))
        else:
            out.append(v)
    return out


def _parse_assignments(body: str) -> List[Tuple[str, str, List[str]]]:
    """[(name, section-or-'' , [value tokens])] in file order."""
    # mask quoted strings so that '=' or names inside them are not matched
    masked = list(body)
    q = None
    for i, ch in enumerate(body):
        if q:
            if ch == q:
                q = None
            else:
                masked[i] = "_"
        elif ch in "'\"":
            q = ch
    masked = "".join(masked)
    ms = list(_ASSIGN_RE.finditer(masked))
    res = []
    for n, m in enumerate(ms):
        end = ms[n + 1].start() if n + 1 < len(ms) else len(body)
        res.append((m.group(1).lower(), (m.group(2) or "").strip(), _split_values(body[m.end():end])))
    return res


def _conv(tok: str, kind: str):
    if kind == "f":
        return float(tok.lower().replace("d", "e"))
    if kind == "i":
        return int(tok)
    if kind == "l":
        t = tok.lower().strip(".")
        if t in ("t", "true"):
            return True
        if t in ("f", "false"):
            return False
        raise InputError(f"bad logical value {tok!r}")
    if kind == "s":
        if len(tok) >= 2 and tok[0] in "'\"" and tok[-1] == tok[0]:
            return tok[1:-1]
        return tok
    raise AssertionError(kind)


class _Var:
    """One namelist variable: a scalar (shape ()) or an array with Fortran lower bounds."""

    def __init__(self, kind: str, shape=(), lower=None, default=None):
        self.kind, self.shape = kind, tuple(shape)
        self.lower = tuple(lower) if lower is not None else tuple(1 for _ in self.shape)
        dtype = {"f": np.float64, "i": np.int64, "l": np.bool_, "s": object}[kind]
        if self.shape:
            self.value = np.empty(self.shape, dtype=dtype, order="F")
            self.value[...] = default
        else:
            self.value = default

    def assign(self, section: str, toks: Sequence[str]):
        vals = [_conv(t, self.kind) for t in toks]
        if not self.shape:
            if section:
                raise InputError("section on a scalar")
            if not vals:
                raise InputError("missing value")
            self.value = vals[0]
            return
        ranges = []
        if section:
            parts = [p.strip() for p in section.strip("()").split(",")]
            if len(parts) != len(self.shape):
                raise InputError(f"rank mismatch in section {section}")
            for p, lo, n in zip(parts, self.lower, self.shape):
                if ":" in p:
                    a, b = p.split(":")[:2]
                    a = int(a) if a.strip() else lo
                    b = int(b) if b.strip() else lo + n - 1
                    ranges.append(range(a - lo, b - lo + 1))
                else:
                    ranges.append(range(int(p) - lo, int(p) - lo + 1))
        else:
            ranges = [range(n) for n in self.shape]
        for r, n in zip(ranges, self.shape):
            if len(r) and (r[0] < 0 or r[-1] >= n):
                raise InputError(f"section {section} out of bounds")
        # column-major: first index fastest; assign only as many elements as values given
        it = itertools.product(*[list(r) for r in reversed(ranges)])
        for v, idx in zip(vals, it):
            self.value[tuple(reversed(idx))] = v


def _read_group(text: str, group: str, schema: Dict[str, _Var], required: bool) -> bool:
    body = _find_group(text, group)
    if body is None:
        if required:
            raise InputError(f"Error reading {group} namelist: group not found")
        return False
    for name, section, toks in _parse_assignments(body):
        if name not in schema:
            raise InputError(f"Error reading {group} namelist: unknown variable {name!r}")
        try:
            schema[name].assign(section, toks)
        except (ValueError, InputError) as e:
            raise InputError(f"Error reading {group} namelist: {name}{section}: {e}") from e
    return True


# --- parameter containers -------------------------------------------------------------------
@dataclasses.dataclass
class BlockParams:
    """Per-block input (reference src/param.f90:46-57). Index conventions follow Fortran:
    ``cbcvel[ib, idir, ivel]`` = ``cbcvel(ib,idir+1,ivel+1)``; ``cbcpre[ib, idir]``."""
    dims: np.ndarray
    ng: np.ndarray
    lmin: np.ndarray
    lmax: np.ndarray
    gt: np.ndarray
    gr: np.ndarray
    cbcvel: np.ndarray   # (2,3,3) of 1-char str
    bcvel: np.ndarray    # (2,3,3)
    cbcpre: np.ndarray   # (2,3)
    bcpre: np.ndarray    # (2,3)
    cbcscal: np.ndarray  # (2,3,nscal)
    bcscal: np.ndarray   # (2,3,nscal)
    inflow_type: np.ndarray  # (2,3)
    inivel: str


@dataclasses.dataclass
class Params:
    """Global input (reference src/param.f90:23-45) + build-time switches made run-time."""
    cfl: float = 0.95
    dtmax: float = 1.0e5
    dt_f: float = -1.0
    visci: float = 1.0
    nstep: int = 0
    time_max: float = 0.0
    tw_max: float = 0.0
    stop_type: Tuple[bool, bool, bool] = (True, False, False)
    restart: bool = False
    is_overwrite_save: bool = True
    nsaves_max: int = 0
    icheck: int = 0
    iout0d: int = 0
    iout1d: int = 0
    iout2d: int = 0
    iout3d: int = 0
    isave: int = 0
    bforce: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))
    gacc: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(3))
    nscal: int = 0
    alphai: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))
    beta: float = 0.0
    iniscal: List[str] = dataclasses.field(default_factory=list)
    ssource: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))
    is_sforced: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, bool))
    scalf: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0))
    alpha_max: float = 0.0
    hypre_solver_i: int = HYPRE_SOLVER_I_DEFAULT
    hypre_tol: float = HYPRE_TOL_DEFAULT
    hypre_maxiter: int = HYPRE_MAXITER_DEFAULT
    nblocks: int = 0
    blocks: List[BlockParams] = dataclasses.field(default_factory=list)
    # build-time switches of the reference (Makefile FFT_AXIS/IMPDIFF/BOUSSINESQ_BUOYANCY)
    fft_axis: int = 0
    impdiff: bool = False
    boussinesq: bool = False

    @property
    def visc(self) -> float:
        return self.visci ** (-1)

    @property
    def nrank(self) -> int:
        return int(sum(int(np.prod(b.dims)) for b in self.blocks))

    @property
    def is_periodic(self) -> np.ndarray:
        """All blocks 'FF' in a direction (reference src/param.f90:206,229-234)."""
        per = np.ones(3, bool)
        for b in self.blocks:
            for q in range(3):
                per[q] &= (b.cbcpre[0, q] + b.cbcpre[1, q]) == "FF"
        return per

    @property
    def vol_all(self) -> float:
        """Sum over ranks of product(lmax-lmin)/product(dims) (reference src/param.f90:239-240)."""
        return float(sum(np.prod(b.lmax - b.lmin) / np.prod(b.dims) * int(np.prod(b.dims))
                         for b in self.blocks))


def fft_axes_allowed(p: "Params") -> List[int]:
    """The FFT_AXIS builds of the reference that accept this input, as 1-based axes (``chk_grid_fft``,
    src/sanity.f90:66-100: the direction is not decomposed, has the same extent and number of points in every block and
    a uniform grid; ``find_fft`` / ``eigenvalues``, src/fft.f90:96-189: the pressure boundary-condition pair has a
    transform and eigenvalues -- 'DN' has none).  With implicit diffusion the velocity and scalar pairs must have them too."""
    out = []
    for d in range(3):
        ok = len(p.blocks) > 0
        b0 = p.blocks[0] if p.blocks else None
        pairs = set()
        for b in p.blocks:
            ok = ok and int(b.dims[d]) == 1 and float(b.gr[d]) == 0.0
            ok = ok and int(b.ng[d]) == int(b0.ng[d]) and float(b.lmin[d]) == float(b0.lmin[d]) and float(b.lmax[d]) == float(b0.lmax[d])
            pairs.add(str(b.cbcpre[0, d]) + str(b.cbcpre[1, d]))
            if p.impdiff:
                for c in range(3):
                    pairs.add(str(b.cbcvel[0, d, c]) + str(b.cbcvel[1, d, c]))
                for i in range(p.nscal):
                    pairs.add(str(b.cbcscal[0, d, i]) + str(b.cbcscal[1, d, i]))
        ok = ok and len(pairs) > 0 and pairs <= {"FF", "NN", "DD", "ND"}
        # every block must carry the same pressure pair along the transform direction (one plan for the run)
        ok = ok and len({str(b.cbcpre[0, d]) + str(b.cbcpre[1, d]) for b in p.blocks}) == 1
        if ok:
            out.append(d + 1)
    return out


def pick_fft_axis(p: "Params") -> int:
    """``--fft-axis auto``: FFT_AXIS is a build option of the reference and a run-time field here; when the input does not
    say, take a build the reference itself would accept for this case (the transform path replaces a 3-D solve by
    per-wavenumber plane solves, direct where the planes are separable), z before y before x; 0 if none applies."""
    ok = fft_axes_allowed(p)
    for d in (3, 2, 1):
        if d in ok:
            return d
    return 0


def read_input(rundir: str = ".", fft_axis: int = 0, impdiff: bool = False,
               boussinesq: bool = False, dns_file: str = "dns.nml",
               blocks_file: str = "blocks.nml") -> Params:
    """Read ``dns.nml`` and ``blocks.nml`` from ``rundir`` (reference ``read_input``,
    src/param.f90:68-277)."""
    path = os.path.join(rundir, dns_file)
    try:
        with open(path) as f:
            text = _strip_comments(f.read())
    except OSError as e:
        raise InputError(f"Error reading dns.nml: {e}") from e
    dns = {
        "cfl": _Var("f", default=0.0), "dtmax": _Var("f", default=0.0), "dt_f": _Var("f", default=-1.0),
        "visci": _Var("f", default=1.0),
        "nstep": _Var("i", default=0), "time_max": _Var("f", default=0.0), "tw_max": _Var("f", default=0.0),
        "stop_type": _Var("l", (3,), default=False),
        "restart": _Var("l", default=False), "is_overwrite_save": _Var("l", default=True),
        "nsaves_max": _Var("i", default=0),
        "icheck": _Var("i", default=0), "iout0d": _Var("i", default=0), "iout1d": _Var("i", default=0),
        "iout2d": _Var("i", default=0), "iout3d": _Var("i", default=0), "isave": _Var("i", default=0),
        "bforce": _Var("f", (3,), default=0.0), "gacc": _Var("f", (3,), default=0.0),
        "nscal": _Var("i", default=0),
    }
    _read_group(text, "dns", dns, required=True)
    hyp = {"hypre_solver_i": _Var("i", default=HYPRE_SOLVER_I_DEFAULT),
           "hypre_tol": _Var("f", default=HYPRE_TOL_DEFAULT),
           "hypre_maxiter": _Var("i", default=HYPRE_MAXITER_DEFAULT)}
    _read_group(text, "hypre", hyp, required=False)
    nscal = max(int(dns["nscal"].value), 0)
    p = Params(fft_axis=fft_axis, impdiff=impdiff, boussinesq=boussinesq)
    for k in ("cfl", "dtmax", "dt_f", "visci", "nstep", "time_max", "tw_max", "restart",
              "is_overwrite_save", "nsaves_max", "icheck", "iout0d", "iout1d", "iout2d", "iout3d", "isave"):
        setattr(p, k, dns[k].value)
    p.stop_type = tuple(bool(x) for x in dns["stop_type"].value)
    p.bforce = np.array(dns["bforce"].value, dtype=np.float64)
    p.gacc = np.array(dns["gacc"].value, dtype=np.float64)
    p.nscal = nscal
    if nscal > 0:
        sc = {"alphai": _Var("f", (nscal,), default=float(np.finfo(np.float64).max)),
              "beta": _Var("f", default=0.0),
              "iniscal": _Var("s", (nscal,), default="zer"),
              "ssource": _Var("f", (nscal,), default=0.0),
              "is_sforced": _Var("l", (nscal,), default=False),
              "scalf": _Var("f", (nscal,), default=0.0)}
        _read_group(text, "scalar", sc, required=True)
        p.alphai = np.array(sc["alphai"].value, dtype=np.float64)
        p.beta = float(sc["beta"].value)
        p.iniscal = [str(s) for s in sc["iniscal"].value]
        p.ssource = np.array(sc["ssource"].value, dtype=np.float64)
        p.is_sforced = np.array(sc["is_sforced"].value, dtype=bool)
        p.scalf = np.array(sc["scalf"].value, dtype=np.float64)
        p.alpha_max = float(np.max(p.alphai ** (-1)))
    if boussinesq and nscal == 0:
        raise InputError("*** Error: _BOUSSINESQ_BUOYANCY requires nscal > 0. ***")
    p.hypre_solver_i = int(hyp["hypre_solver_i"].value)
    p.hypre_tol = float(hyp["hypre_tol"].value)
    p.hypre_maxiter = int(hyp["hypre_maxiter"].value)

    path = os.path.join(rundir, blocks_file)
    try:
        with open(path) as f:
            btext = _strip_comments(f.read())
    except OSError as e:
        raise InputError(f"Error reading blocks.nml: {e}") from e
    nb_ = MAX_BLOCKS
    ns_ = max(nscal, 1)
    blk = {
        "nblocks": _Var("i", default=0),
        "block_dims": _Var("i", (3, nb_), default=1), "block_ng": _Var("i", (3, nb_), default=1),
        "block_lmin": _Var("f", (3, nb_), default=0.0), "block_lmax": _Var("f", (3, nb_), default=1.0),
        "block_gt": _Var("i", (3, nb_), default=0), "block_gr": _Var("f", (3, nb_), default=0.0),
        "block_cbcvel": _Var("s", (2, 3, 3, nb_), (0, 1, 1, 1), default="D"),
        "block_bcvel": _Var("f", (2, 3, 3, nb_), (0, 1, 1, 1), default=0.0),
        "block_cbcpre": _Var("s", (2, 3, nb_), (0, 1, 1), default="N"),
        "block_bcpre": _Var("f", (2, 3, nb_), (0, 1, 1), default=0.0),
        "block_cbcscal": _Var("s", (2, 3, ns_, nb_), (0, 1, 1, 1), default="N"),
        "block_bcscal": _Var("f", (2, 3, ns_, nb_), (0, 1, 1, 1), default=0.0),
        "block_inflow_type": _Var("i", (2, 3, nb_), (0, 1, 1), default=0),
        "block_inivel": _Var("s", (nb_,), default="zer"),
    }
    _read_group(btext, "blocks", blk, required=True)
    nblocks = int(blk["nblocks"].value)
    if nblocks < 1 or nblocks > MAX_BLOCKS:
        raise InputError("*** Error: invalid number of blocks in blocks.nml ***")
    p.nblocks = nblocks
    for ib in range(nblocks):
        p.blocks.append(BlockParams(
            dims=np.array(blk["block_dims"].value[:, ib], dtype=np.int64),
            ng=np.array(blk["block_ng"].value[:, ib], dtype=np.int64),
            lmin=np.array(blk["block_lmin"].value[:, ib], dtype=np.float64),
            lmax=np.array(blk["block_lmax"].value[:, ib], dtype=np.float64),
            gt=np.array(blk["block_gt"].value[:, ib], dtype=np.int64),
            gr=np.array(blk["block_gr"].value[:, ib], dtype=np.float64),
            cbcvel=np.array(blk["block_cbcvel"].value[:, :, :, ib], dtype=object),
            bcvel=np.array(blk["block_bcvel"].value[:, :, :, ib], dtype=np.float64),
            cbcpre=np.array(blk["block_cbcpre"].value[:, :, ib], dtype=object),
            bcpre=np.array(blk["block_bcpre"].value[:, :, ib], dtype=np.float64),
            cbcscal=np.array(blk["block_cbcscal"].value[:, :, :nscal, ib], dtype=object),
            bcscal=np.array(blk["block_bcscal"].value[:, :, :nscal, ib], dtype=np.float64),
            inflow_type=np.array(blk["block_inflow_type"].value[:, :, ib], dtype=np.int64),
            inivel=str(blk["block_inivel"].value[ib]).strip(),
        ))
    # validation of solver parameters (reference src/param.f90:253-268): revert to defaults
    if p.hypre_solver_i < 1 or p.hypre_solver_i > 4:
        p.hypre_solver_i = HYPRE_SOLVER_I_DEFAULT
    if p.hypre_tol > 1.0 or p.hypre_tol < 0.0:
        p.hypre_tol = HYPRE_TOL_DEFAULT
    if p.hypre_maxiter < 0:
        p.hypre_maxiter = HYPRE_MAXITER_DEFAULT
    return p
