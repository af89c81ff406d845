"""fluca_b200 -- B200-native Navier-Stokes time step behind thecasterian/fluca's NS type interface.

csrc/   hand-written CUDA (sm_100a) + the C ABI of include/fluca_b200.h  -> libfluca_b200.so
ns.py   host-side mirror of the reference's NS / MeshCart API for this path (type name "b200")
"""
from . import _lib  # noqa: F401
from .solver import Comm, Solver, slab_partition  # noqa: F401
from .ns import *  # noqa: F401,F403
from . import fd  # noqa: F401
