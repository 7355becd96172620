"""ctypes binding of ``libsnacb200.so`` (C ABI declared in ``include/snac_b200.h``).

This is the stub a host program writes against the C ABI; the Fortran equivalent (``bind(C)``
interface module) is shown in ``INTEGRATION.md``.  There is no CPU fallback: if the shared library
is missing or no CUDA device exists, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SNAC_B200_LIB: an alternative build of the same library (kernel-tuning experiments, tools/build_variant.py)
LIB_PATH = os.environ.get("SNAC_B200_LIB") or os.path.join(HERE, "libsnacb200.so")

SNAC_ABI_VERSION = 2
SNAC_MAX_SCAL = 8

FIELD_U, FIELD_V, FIELD_W, FIELD_P, FIELD_PP, FIELD_RHS, FIELD_PO = range(7)
FIELD_DUDTRKO, FIELD_DVDTRKO, FIELD_DWDTRKO = 7, 8, 9
FIELD_RHS_HAT = 10
FIELD_S0 = 16
FIELD_DSDTRKO0 = 32


class SnacError(RuntimeError):
    pass


class snac_params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("device", C.c_int), ("nrank", C.c_int), ("nproc", C.c_int),
        ("myproc", C.c_int), ("rank_owner", C.POINTER(C.c_int)), ("periods", C.c_int * 3),
        ("fft_axis", C.c_int), ("impdiff", C.c_int), ("boussinesq", C.c_int), ("nscal", C.c_int),
        ("visc", C.c_double), ("bforce", C.c_double * 3), ("gacc", C.c_double * 3), ("beta", C.c_double),
        ("vol_all", C.c_double), ("alphai", C.c_double * SNAC_MAX_SCAL), ("ssource", C.c_double * SNAC_MAX_SCAL),
        ("is_sforced", C.c_int * SNAC_MAX_SCAL), ("scalf", C.c_double * SNAC_MAX_SCAL),
        ("hypre_tol", C.c_double), ("hypre_maxiter", C.c_int), ("hypre_solver_i", C.c_int),
        ("mg_npre", C.c_int), ("mg_npost", C.c_int), ("mg_nbottom", C.c_int), ("mg_generic", C.c_int),
        ("plane_solver", C.c_int),
    ]


class snac_box(C.Structure):
    _fields_ = [
        ("rank", C.c_int), ("block", C.c_int), ("lo", C.c_int * 3), ("hi", C.c_int * 3),
        ("nb", (C.c_int * 2) * 3), ("is_bound", (C.c_int * 2) * 3),
        ("cbcvel", ((C.c_char * 2) * 3) * 3), ("bcvel", ((C.c_double * 2) * 3) * 3),
        ("cbcpre", (C.c_char * 2) * 3), ("bcpre", (C.c_double * 2) * 3),
        ("cbcscal", ((C.c_char * 2) * 3) * SNAC_MAX_SCAL), ("bcscal", ((C.c_double * 2) * 3) * SNAC_MAX_SCAL),
        ("dxc", C.POINTER(C.c_double)), ("dxf", C.POINTER(C.c_double)),
        ("dyc", C.POINTER(C.c_double)), ("dyf", C.POINTER(C.c_double)),
        ("dzc", C.POINTER(C.c_double)), ("dzf", C.POINTER(C.c_double)),
        ("dl", (C.c_double * 2) * 3), ("dlf", (C.c_double * 2) * 3),
    ]


EXPORTS = {
    "snac_create": (C.c_int, [C.POINTER(snac_params), C.POINTER(C.c_void_p)]),
    "snac_add_box": (C.c_int, [C.c_void_p, C.POINTER(snac_box)]),
    "snac_commit": (C.c_int, [C.c_void_p]),
    "snac_comm_blob_size": (C.c_size_t, []),
    "snac_comm_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "snac_comm_connect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "snac_finalize": (C.c_int, [C.c_void_p]),
    "snac_last_error": (C.c_char_p, [C.c_void_p]),
    "snac_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "snac_download": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "snac_download_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int)]),
    "snac_download_wait": (C.c_int, [C.c_void_p, C.c_int]),
    "snac_host_alloc": (C.c_void_p, [C.c_size_t]),
    "snac_host_free": (None, [C.c_void_p]),
    "snac_rk_mom": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "snac_rk_scal": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_int]),
    "snac_bounduvw": (C.c_int, [C.c_void_p, C.c_int]),
    "snac_boundp": (C.c_int, [C.c_void_p, C.c_int]),
    "snac_set_inflow": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snac_inflow": (C.c_int, [C.c_void_p, C.c_int]),
    "snac_fillps": (C.c_int, [C.c_void_p, C.c_double]),
    "snac_updt_rhs": (C.c_int, [C.c_void_p]),
    "snac_fft": (C.c_int, [C.c_void_p, C.c_int]),
    "snac_solve_n_helmholtz_2d": (C.c_int, [C.c_void_p]),
    "snac_solve_helmholtz": (C.c_int, [C.c_void_p]),
    "snac_correc": (C.c_int, [C.c_void_p, C.c_double]),
    "snac_updt_pressure": (C.c_int, [C.c_void_p, C.c_double]),
    "snac_chkdt": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "snac_chkdiv": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "snac_out1d": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]),
    "snac_out2d": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p]),
    "snac_step": (C.c_int, [C.c_void_p, C.c_double]),
    "snac_solver_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "snac_solver_mean_iters": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "snac_plane_solver": (C.c_int, [C.c_void_p]),
    "snac_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "snac_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "snac_profile_units": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double)]),
    "snac_profile_name": (C.c_char_p, [C.c_int]),
    "snac_timer_start": (C.c_int, [C.c_void_p]),
    "snac_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "snac_launch_count": (C.c_longlong, [C.c_void_p]),
    "snac_sync": (C.c_int, [C.c_void_p]),
    "snac_stream": (C.c_void_p, [C.c_void_p]),
}

_lib = None


def load_library(path: str = LIB_PATH):
    """Loads the shared library and declares every entry point of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise SnacError(f"{path} is missing: build it with `python -m snac_b200.build` "
                        "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
