"""The C-ABI library loads and exports every symbol include/fluca_b200.h declares (no compute calls
without a GPU), and the product path has no CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest

import fluca_b200 as fb
from tests import parity

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fluca_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(fluca_b200_[a-z0-9_]+)\s*\(", src))
    names -= {n for n in names if n.endswith("_fn")}
    return sorted(names)


def test_header_and_python_prototypes_agree():
    assert declared_symbols() == sorted(fb._lib.SYMBOLS)


def test_product_library_exports_every_declared_symbol():
    subprocess.run(["make", "-C", os.path.join(ROOT, "fluca_b200", "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    out = subprocess.run(["nm", "-D", "--defined-only", fb._lib.PRODUCT_LIB], check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (fluca_b200_[a-z0-9_]+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, missing
    # it is a real sm_100a build: SASS for sm_100a is embedded
    lst = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "--list-elf", fb._lib.PRODUCT_LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in lst


def test_product_library_loads_and_is_not_the_test_double():
    L = fb._lib.load()
    assert L.fluca_b200_is_host_emulation() == 0
    for s in fb._lib.SYMBOLS:
        assert hasattr(L, s)


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np

    with pytest.raises(fb._lib.FlucaB200Error) as ei:
        fb.Solver((8, 8), [np.linspace(0, 1, 9)] * 2, [1, 1, 1, 1], 1.0, 0.01, 0.01)
    assert ei.value.code == fb._lib.ERR_NODEVICE
    assert "no CPU fallback" in str(ei.value)


def test_host_emulation_double_is_flagged():
    L = parity.hostemu_library()
    assert L.fluca_b200_is_host_emulation() == 1
    # and it cannot be reached through the product loader
    assert os.path.abspath(parity.HOSTEMU) != os.path.abspath(fb._lib.PRODUCT_LIB)


def test_error_codes_and_messages():
    import numpy as np

    L = parity.hostemu_library()
    with pytest.raises(fb._lib.FlucaB200Error) as ei:
        fb.Solver((8, 2), [np.linspace(0, 1, 9), np.linspace(0, 1, 3)], [1, 1, 1, 1], 1.0, 0.01, 0.01, library=L)
    assert ei.value.code == fb._lib.ERR_ARG and "at least 3 cells" in str(ei.value)
    with pytest.raises(fb._lib.FlucaB200Error):
        fb.Solver((8, 8), [np.linspace(0, 1, 9)] * 2, [3, 1, 1, 1], 1.0, 0.01, 0.01, library=L)  # periodic on one side only
    with pytest.raises(fb._lib.FlucaB200Error):
        fb.Solver((8, 8), [np.linspace(0, 1, 9)] * 2, [1, 1, 1, 1], 0.0, 0.01, 0.01, library=L)  # rho = 0
