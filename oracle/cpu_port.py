"""ctypes front end of ``oracle/snac_cpu.c`` (C + OpenMP restatement of the pipeline for the
triperiodic benchmark workload).  TEST / BASELINE INFRASTRUCTURE ONLY -- see the header of the C file.
The shared object is built into ``oracle/_ref/`` (git-ignored, travels to the GPU box)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "snac_cpu.c")
LIB = os.path.join(HERE, "_ref", "libsnac_cpu.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        raise RuntimeError("no C compiler: the CPU baseline cannot be built")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # -march=x86-64-v3 (AVX2/FMA) rather than native: the object is built here and run on the GPU box
    cmd = [cc, "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"building {SRC} failed:\n{r.stdout}\n{r.stderr}")
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.snac_cpu_create.restype = C.c_void_p
        L.snac_cpu_create.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_double, C.c_int]
        L.snac_cpu_init_bc.argtypes = [C.c_void_p]
        L.snac_cpu_set_threads.argtypes = [C.c_int]
        L.snac_cpu_destroy.argtypes = [C.c_void_p]
        L.snac_cpu_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.snac_cpu_get.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.snac_cpu_step.argtypes = [C.c_void_p, C.c_double]
        L.snac_cpu_chkdt.restype = C.c_double
        L.snac_cpu_chkdt.argtypes = [C.c_void_p]
        L.snac_cpu_divmax.restype = C.c_double
        L.snac_cpu_divmax.argtypes = [C.c_void_p]
        L.snac_cpu_mean_cycles.restype = C.c_double
        L.snac_cpu_mean_cycles.argtypes = [C.c_void_p]
        L.snac_cpu_threads.restype = C.c_int
        _lib = L
    return _lib


def use_all_cores():
    """All host cores for the OpenMP loops, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would turn the baseline into a single-thread run)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    _load().snac_cpu_set_threads(int(n))
    return int(n)


class CpuSim:
    """One rectangular domain on a uniform grid; ``walls[d]``: no-slip walls (else periodic) in
    direction d; transforms along ``fft_axis`` (1, 2 or 3: a periodic direction)."""

    def __init__(self, n, lengths, visc, tol=1e-4, maxiter=50, walls=(0, 0, 0), fft_axis=3, bforce=(0.0, 0.0, 0.0)):
        self.L = _load()
        self.n = tuple(int(x) for x in n)
        ni = (C.c_int * 3)(*self.n)
        le = (C.c_double * 3)(*[float(x) for x in lengths])
        wa = (C.c_int * 3)(*[int(bool(x)) for x in walls])
        bf = (C.c_double * 3)(*[float(x) for x in bforce])
        self.h = self.L.snac_cpu_create(ni, le, wa, int(fft_axis), float(visc), bf, float(tol), int(maxiter))
        if not self.h:
            raise ValueError("snac_cpu_create: the transform axis must be a periodic direction (or out of memory)")
        self.shape = tuple(x + 2 for x in self.n)

    def init_bc(self):
        self.L.snac_cpu_init_bc(self.h)

    @property
    def threads(self):
        return int(self.L.snac_cpu_threads())

    def set(self, which, a):
        a = np.asfortranarray(a, dtype=np.float64)
        assert a.shape == self.shape
        self.L.snac_cpu_set(self.h, which, a.ctypes.data_as(C.c_void_p))

    def get(self, which):
        a = np.empty(self.shape, dtype=np.float64, order="F")
        self.L.snac_cpu_get(self.h, which, a.ctypes.data_as(C.c_void_p))
        return a

    def step(self, dt):
        self.L.snac_cpu_step(self.h, float(dt))

    def chkdt(self):
        return float(self.L.snac_cpu_chkdt(self.h))

    def divmax(self):
        return float(self.L.snac_cpu_divmax(self.h))

    def mean_cycles(self):
        return float(self.L.snac_cpu_mean_cycles(self.h))

    def close(self):
        if self.h:
            self.L.snac_cpu_destroy(self.h)
            self.h = None
