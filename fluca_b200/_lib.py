"""Loader and ctypes prototypes of the C ABI (include/fluca_b200.h).

The product library is fluca_b200/csrc/libfluca_b200.so (nvcc, sm_100a), built in-tree by
__graft_entry__.build().  There is NO CPU fallback: if the library is missing, or no CUDA device is
visible, loading / solver creation fails loudly.  The host-emulation test double
(tests/hostemu/_build/libfluca_b200_hostemu.so) can only be selected explicitly by the CPU tests
through load(path=...); nothing in this package ever picks it.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.path.join(_HERE, "csrc", "libfluca_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_NCCL, ERR_DIVERGED, ERR_NODEVICE, ERR_INTERNAL = range(7)


class FlucaB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"fluca_b200 error {code}: {msg}")
        self.code = code


class Desc(C.Structure):
    _fields_ = [
        ("dim", C.c_int),
        ("n", C.c_int * 3),
        ("xf", C.POINTER(C.c_double) * 3),
        ("bc_type", C.c_int * 6),
        ("rho", C.c_double),
        ("mu", C.c_double),
        ("dt", C.c_double),
        ("k0", C.c_int),
        ("nzl", C.c_int),
        ("mode", C.c_int),
        ("outer_rtol", C.c_double),
        ("outer_maxit", C.c_int),
        ("outer_restart", C.c_int),
        ("mom_rtol", C.c_double),
        ("schur_rtol", C.c_double),
        ("inner_maxit", C.c_int),
        ("mg_nu1", C.c_int),
        ("mg_nu2", C.c_int),
        ("mg_coarse_sweeps", C.c_int),
        ("no_bcg_quirk", C.c_int),
        ("no_t_outlet_quirk", C.c_int),
    ]


INNER_MONITOR_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_double)  # fluca_b200_inner_monitor_fn


class Stats(C.Structure):
    _fields_ = [
        ("outer_its", C.c_int),
        ("mom_its", C.c_int),
        ("schur_its", C.c_int),
        ("abf_applies", C.c_int),
        ("converged", C.c_int),
        ("outer_rnorm0", C.c_double),
        ("outer_rnorm", C.c_double),
        ("nhist", C.c_int),
        ("hist", C.c_double * 128),
        ("launches", C.c_long),
        ("mom_last_rel", C.c_double),
        ("schur_last_rel", C.c_double),
        ("inner_unconverged", C.c_int),
    ]


HALO_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long, C.c_int)
ALLSUM_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.c_int)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_long)

# every symbol include/fluca_b200.h declares (tests/test_abi.py checks the export list against the header)
SYMBOLS = [
    "fluca_b200_last_error",
    "fluca_b200_is_host_emulation",
    "fluca_b200_comm_unique_id",
    "fluca_b200_comm_create_nccl",
    "fluca_b200_comm_create_callbacks",
    "fluca_b200_create",
    "fluca_b200_destroy",
    "fluca_b200_set_state",
    "fluca_b200_get_state",
    "fluca_b200_set_boundary_velocity",
    "fluca_b200_set_boundary_pressure",
    "fluca_b200_step",
    "fluca_b200_prepare_step",
    "fluca_b200_get_rhs",
    "fluca_b200_apply_momentum",
    "fluca_b200_apply_schur",
    "fluca_b200_apply_vcycle",
    "fluca_b200_apply_coupled",
    "fluca_b200_apply_abf",
    "fluca_b200_snapshot_save",
    "fluca_b200_snapshot_restore",
    "fluca_b200_device_layout",
    "fluca_b200_device_field",
    "fluca_b200_stream",
    "fluca_b200_launch_count",
    "fluca_b200_step_model_bytes",
    "fluca_b200_kernel_timing",
    "fluca_b200_kernel_times",
    "fluca_b200_time_kernel",
    "fluca_b200_set_markers",
    "fluca_b200_get_marker_forces",
    "fluca_b200_ibm_interpolate",
    "fluca_b200_ibm_spread",
    "fluca_b200_set_ibm_iterations",
    "fluca_b200_set_abf_ainv_types",
    "fluca_b200_set_inner_monitor",
    "fluca_b200_stage_state",
    "fluca_b200_staged_state",
    # FlucaFD stencil layer (host-side; csrc/fd.cu)
    "fluca_b200_fd_last_error",
    "fluca_b200_fd_grid_create",
    "fluca_b200_fd_grid_destroy",
    "fluca_b200_fd_derivative_create",
    "fluca_b200_fd_sum_create",
    "fluca_b200_fd_scale_create_constant",
    "fluca_b200_fd_scale_create_vector",
    "fluca_b200_fd_composition_create",
    "fluca_b200_fd_tvd_create",
    "fluca_b200_fd_tvd_set_limiter",
    "fluca_b200_fd_tvd_set_velocity",
    "fluca_b200_fd_tvd_set_current_solution",
    "fluca_b200_fd_set_locations",
    "fluca_b200_step_model_bytes_split",
    "fluca_b200_ibm_info",
    "fluca_b200_host_alloc",
    "fluca_b200_host_free",
    "fluca_b200_fd_set_boundary_condition",
    "fluca_b200_fd_setup",
    "fluca_b200_fd_get_stencil",
    "fluca_b200_fd_apply_inputs",
    "fluca_b200_fd_apply",
    "fluca_b200_fd_apply_device",
    "fluca_b200_fd_stream",
    "fluca_b200_fd_sync",
    "fluca_b200_fd_get_operator",
    "fluca_b200_fd_destroy",
]
FD_MAX_STENCIL = 32


class FDCol(C.Structure):
    _fields_ = [("i", C.c_int), ("j", C.c_int), ("k", C.c_int), ("loc", C.c_int), ("c", C.c_int)]
# PCABFAinvType (flucans.h:99-103) and its option strings PCABFAinvTypes[] (abfpc.c)
AINV_ID, AINV_DIAG, AINV_ROWSUM = 0, 1, 2
AINV_NAMES = {"id": AINV_ID, "diag": AINV_DIAG, "rowsum": AINV_ROWSUM}
KT_NAMES = ["momentum_apply", "momentum_vec", "poisson_apply", "poisson_vec", "mg_smooth", "mg_transfer", "rhs_project", "outer", "halo", "ibm"]

_P = C.c_void_p
_PD3 = C.POINTER(C.c_void_p)


def _prototype(L):
    L.fluca_b200_last_error.restype = C.c_char_p
    L.fluca_b200_is_host_emulation.restype = C.c_int
    L.fluca_b200_comm_unique_id.argtypes = [_P, C.c_int, C.POINTER(C.c_int)]
    L.fluca_b200_comm_create_nccl.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]
    L.fluca_b200_comm_create_callbacks.argtypes = [C.c_int, C.c_int, HALO_FN, ALLSUM_FN, ALLGATHER_FN, _P, C.POINTER(_P)]
    L.fluca_b200_create.argtypes = [C.POINTER(Desc), _P, C.POINTER(_P)]
    L.fluca_b200_destroy.argtypes = [_P]
    L.fluca_b200_set_state.argtypes = [_P, _P, _PD3, _P, _P]
    L.fluca_b200_get_state.argtypes = [_P, _P, _PD3, _P, _P]
    L.fluca_b200_set_boundary_velocity.argtypes = [_P, C.c_int, C.c_int, _P]
    L.fluca_b200_set_boundary_pressure.argtypes = [_P, C.c_int, C.c_int, _P]
    L.fluca_b200_step.argtypes = [_P, C.c_double, C.c_int, C.POINTER(Stats)]
    L.fluca_b200_prepare_step.argtypes = [_P, C.c_double, C.c_int]
    L.fluca_b200_get_rhs.argtypes = [_P, _P, _PD3, _P]
    L.fluca_b200_apply_momentum.argtypes = [_P, _P, _P]
    L.fluca_b200_apply_schur.argtypes = [_P, _P, _P]
    L.fluca_b200_apply_vcycle.argtypes = [_P, _P, _P]
    L.fluca_b200_apply_coupled.argtypes = [_P, _P, _PD3, _P, _P, _PD3, _P]
    L.fluca_b200_apply_abf.argtypes = [_P, _P, _PD3, _P, _P, _PD3, _P, C.POINTER(Stats)]
    L.fluca_b200_snapshot_save.argtypes = [_P]
    L.fluca_b200_snapshot_restore.argtypes = [_P]
    L.fluca_b200_device_layout.argtypes = [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long)]
    L.fluca_b200_device_field.argtypes = [_P, C.c_char_p, C.POINTER(_P)]
    L.fluca_b200_stream.argtypes = [_P, C.POINTER(_P)]
    L.fluca_b200_launch_count.argtypes = [_P]
    L.fluca_b200_launch_count.restype = C.c_long
    L.fluca_b200_step_model_bytes.argtypes = [_P, C.POINTER(Stats)]
    L.fluca_b200_step_model_bytes.restype = C.c_double
    L.fluca_b200_step_model_bytes_split.argtypes = [_P, C.POINTER(Stats), C.POINTER(C.c_double)]
    L.fluca_b200_kernel_timing.argtypes = [_P, C.c_int]
    L.fluca_b200_kernel_times.argtypes = [_P, C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_int]
    L.fluca_b200_time_kernel.argtypes = [_P, C.c_char_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.fluca_b200_set_markers.argtypes = [_P, C.c_long, _P, _P, _P, C.c_int]
    L.fluca_b200_get_marker_forces.argtypes = [_P, _P, _P]
    L.fluca_b200_ibm_info.argtypes = [_P, C.POINTER(C.c_long)]
    L.fluca_b200_host_alloc.argtypes = [C.c_size_t, C.POINTER(_P)]
    L.fluca_b200_host_free.argtypes = [_P]
    L.fluca_b200_ibm_interpolate.argtypes = [_P, _P, _P]
    L.fluca_b200_ibm_spread.argtypes = [_P, _P, _P]
    L.fluca_b200_set_ibm_iterations.argtypes = [_P, C.c_int]
    L.fluca_b200_set_abf_ainv_types.argtypes = [_P, C.c_int, C.c_int]
    L.fluca_b200_set_inner_monitor.argtypes = [_P, INNER_MONITOR_FN, C.c_void_p]
    L.fluca_b200_fd_last_error.restype = C.c_char_p
    L.fluca_b200_fd_grid_create.argtypes = [C.c_int, C.POINTER(C.c_int), _PD3, _PD3, C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]
    L.fluca_b200_fd_grid_destroy.argtypes = [_P]
    L.fluca_b200_fd_derivative_create.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]
    L.fluca_b200_fd_sum_create.argtypes = [C.c_int, C.POINTER(_P), C.POINTER(_P)]
    L.fluca_b200_fd_scale_create_constant.argtypes = [_P, C.c_double, C.POINTER(_P)]
    L.fluca_b200_fd_scale_create_vector.argtypes = [_P, _P, C.c_int, C.c_int, C.POINTER(_P)]
    L.fluca_b200_fd_composition_create.argtypes = [_P, _P, C.POINTER(_P)]
    L.fluca_b200_fd_tvd_create.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]
    L.fluca_b200_fd_tvd_set_limiter.argtypes = [_P, C.c_char_p]
    L.fluca_b200_fd_tvd_set_velocity.argtypes = [_P, _P]
    L.fluca_b200_fd_tvd_set_current_solution.argtypes = [_P, _P]
    L.fluca_b200_fd_set_locations.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fluca_b200_fd_set_boundary_condition.argtypes = [_P, C.c_int, C.c_int, C.c_double]
    L.fluca_b200_fd_setup.argtypes = [_P]
    L.fluca_b200_fd_get_stencil.argtypes = [_P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(FDCol), C.POINTER(C.c_double)]
    L.fluca_b200_fd_apply_inputs.argtypes = [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.fluca_b200_fd_apply.argtypes = [_P, C.c_int, C.POINTER(_P), _P]
    L.fluca_b200_fd_apply_device.argtypes = [_P, C.c_int, C.POINTER(_P), _P]
    L.fluca_b200_fd_stream.argtypes = [_P, C.POINTER(_P)]
    L.fluca_b200_fd_sync.argtypes = [_P]
    L.fluca_b200_fd_get_operator.argtypes = [_P, C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(FDCol), C.POINTER(C.c_double)]
    L.fluca_b200_fd_destroy.argtypes = [_P]
    L.fluca_b200_stage_state.argtypes = [_P]
    L.fluca_b200_staged_state.argtypes = [_P, C.POINTER(_P), _PD3, C.POINTER(_P), C.POINTER(_P)]
    for name in SYMBOLS:  # every declared symbol must resolve (a stale .so fails here, not at first use)
        getattr(L, name)
    return L


_cache = {}


def load(path: str | None = None):
    """Load the C-ABI library.  path=None -> the product library (CUDA, sm_100a)."""
    path = os.path.abspath(path or PRODUCT_LIB)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise FlucaB200Error(ERR_NODEVICE, f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a). There is no CPU fallback.")
    L = _prototype(C.CDLL(path))
    if path == os.path.abspath(PRODUCT_LIB) and L.fluca_b200_is_host_emulation():
        raise FlucaB200Error(ERR_INTERNAL, "the product library path holds a host-emulation build")
    _cache[path] = L
    return L


def check(L, rc):
    if rc != OK:
        raise FlucaB200Error(rc, (L.fluca_b200_last_error() or b"").decode())
