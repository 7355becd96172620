"""Builds ``libsnacb200.so`` (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

The library has no Python or torch dependency: plain CUDA runtime, linked statically
(``-cudart static``) so that it loads next to any other CUDA user in the process.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsnacb200.so")
SOURCES = ["api.cu", "mom.cu", "mom_tma.cu", "bound.cu", "pois.cu", "pois_tma.cu", "fft.cu", "mg.cu", "mg_ring.cu", "mg_coarse.cu", "mg_coop.cu", "direct.cu", "comm.cu", "scal.cu", "output.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _nvcc() -> str:
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found: libsnacb200 cannot be built")
    return nv


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "snac_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """``defines``/``out``: a tuning variant of the library (extra -D flags) written to ``out``."""
    if out is None and not force and not _stale():
        return LIB
    nv = _nvcc()
    objdir = os.path.join(HERE, "build" if out is None else "build_" + os.path.basename(out).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nv, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, err in res:
            sys.stderr.write(err)
    cmd = [nv, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", out or LIB, *[o for o, _ in res]]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return out or LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
