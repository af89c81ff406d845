"""pytest plugin of tools/gpu_suite_on_hostemu.sh -- TEST TOOLING, never loaded by the product or by the normal test runs.

Runs the BODIES of the `-m gpu` tests on a machine without a GPU: `fluca_b200._lib.load()` hands out the host-emulation build
(tests/hostemu: the same solver sources compiled with g++, kernel functors in serial loops) where a test asks for the product
library, and reports "not host emulation" so that the tests' guards pass.  What it checks is everything the two builds share -- the
step driver, Krylov and multigrid logic, boundary handling, IBM, the C ABI -- on the GPU tests' own cases and thresholds; what it
cannot check is what differs: the TMA tile kernels, grid-level reductions, streams, NCCL.  A green run here is therefore NOT a GPU
result and is never reported as one; it is a pre-flight for changes made when no GPU time is left."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

torch.cuda.is_available = lambda: True  # tests/conftest.py skips -m gpu tests otherwise

from fluca_b200 import _lib  # noqa: E402

HOSTEMU = os.environ.get("FLUCA_B200_HOSTEMU_LIB") or os.path.join(ROOT, "tests", "hostemu", "_build", "libfluca_b200_hostemu.so")
_real_load, _proxy = _lib.load, None


class _NotEmulated:
    def __init__(self, lib):
        object.__setattr__(self, "_lib", lib)

    def __getattr__(self, name):
        if name == "fluca_b200_is_host_emulation":
            return lambda: 0
        return getattr(self._lib, name)


def _load(path=None):
    global _proxy
    if path is None or os.path.abspath(path) == os.path.abspath(_lib.PRODUCT_LIB):
        if _proxy is None:
            _proxy = _NotEmulated(_real_load(HOSTEMU))
        return _proxy
    return _real_load(path)


_lib.load = _load
