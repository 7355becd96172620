"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports
every symbol include/snac_b200.h declares, and refuses to run without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from snac_b200.build import build_library
    from snac_b200 import lib as L
    build_library()
    return L.load_library()


def _declared():
    with open(os.path.join(ROOT, "include", "snac_b200.h")) as f:
        text = f.read()
    return sorted(set(re.findall(r"SNAC_API[^;(]*?\b(snac_\w+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from snac_b200 import lib as L
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in L.EXPORTS, f"{n} has no ctypes prototype in snac_b200/lib.py"
    assert sorted(L.EXPORTS) == names


def test_struct_layout_matches_header_sizes(lib):
    """sizeof of the ctypes mirrors equals what gcc computes from the header."""
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "snac_b200.h"\nint main(){printf("%zu %zu\\n", sizeof(snac_params), sizeof(snac_box));return 0;}\n'
    d = tempfile.mkdtemp()
    with open(os.path.join(d, "s.c"), "w") as f:
        f.write(src)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
    a, b = map(int, subprocess.check_output([os.path.join(d, "s")]).split())
    from snac_b200 import lib as L
    assert C.sizeof(L.snac_params) == a
    assert C.sizeof(L.snac_box) == b


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from snac_b200 import lib as L
    prm = L.snac_params()
    prm.abi_version = L.SNAC_ABI_VERSION
    prm.nrank = prm.nproc = 1
    h = C.c_void_p()
    assert lib.snac_create(C.byref(prm), C.byref(h)) != 0
    assert b"no CUDA device" in lib.snac_last_error(None)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "snac_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".h")):
                with open(os.path.join(dirpath, fn)) as f:
                    assert "oracle" not in f.read().lower().replace("there is no cpu", ""), fn
