"""The register-radix transform kernels (snac_b200/csrc/fft_radix.cuh) run thread by thread on the CPU
(tests/host/fft_emul.cu, compiled with nvcc as host code) against the dense definitions of the FFTW r2r
kinds the reference selects (src/fft.f90:96-149): index algebra, twiddles, split-array addressing."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_radix_kernels_on_the_host(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "fft_emul")
    # the translation unit also holds the device kernels (cp.async, shuffles): they are compiled for the target
    # architecture and never launched here; only the __host__ __device__ phase functions run, on the CPU
    r = subprocess.run([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                        os.path.join(ROOT, "tests", "host", "fft_emul.cu")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.startswith("OK")
