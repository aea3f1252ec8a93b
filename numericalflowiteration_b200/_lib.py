"""ctypes binding of libnufi_b200.so (include/nufi_b200.h).  Fails loudly when the CUDA library is missing:
there is no CPU or PyTorch fallback for the hot path."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NUFI_B200_LIB") or os.path.join(HERE, "lib", "libnufi_b200.so")  # override: tuning experiments only

OK, ERR_RANGE, ERR_CUDA, ERR_ALLOC, ERR_ARG = 0, 1, 2, 3, 4

# every symbol include/nufi_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "nufi_b200_create_1d", "nufi_b200_create_2d", "nufi_b200_create_3d", "nufi_b200_destroy", "nufi_b200_last_error",
    "nufi_b200_compute_rho", "nufi_b200_download_rho", "nufi_b200_upload_phi", "nufi_b200_compute_metrics",
    "nufi_b200_download_metrics", "nufi_b200_eval_rho_all", "nufi_b200_solve_interpolate",
    "nufi_b200_solve_interpolate_host", "nufi_b200_poisson_solve", "nufi_b200_interpolate", "nufi_b200_step", "nufi_b200_step_host", "nufi_b200_download_energy", "nufi_b200_download_phi",
    "nufi_b200_sync", "nufi_b200_download_history", "nufi_b200_upload_history", "nufi_b200_eval_f",
    "nufi_b200_eval_field", "nufi_b200_set_stream", "nufi_b200_rho_device", "nufi_b200_field_tail_device",
    "nufi_b200_launch_count", "nufi_b200_last_backtrace_ms", "nufi_b200_backtrace_time", "nufi_b200_last_variant", "nufi_b200_set_variant",
    "nufi_b200_set_tail_variant", "nufi_b200_last_tail_variant", "nufi_b200_measure_fp64_peak", "nufi_b200_version",
    "nufi_b200_group_create", "nufi_b200_group_destroy", "nufi_b200_group_step", "nufi_b200_group_sync",
    "nufi_b200_group_last_error", "nufi_b200_device_count", "nufi_b200_device_of",
    "nufi_b200_peer_export", "nufi_b200_peer_attach", "nufi_b200_peer_step", "nufi_b200_peer_status", "nufi_b200_peer_detach",
    "nufi_b200_group_set_exchange", "nufi_b200_group_exchange", "nufi_b200_set_kernel_timing", "nufi_b200_set_tile_nodes", "nufi_b200_eval_phase_flow",
    "nufi_b200_set_metrics_grid_1d", "nufi_b200_download_rho_full",
]

_lib = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(HERE, "csrc"), "-j4"] + ([] if verbose else ["-s"])
    subprocess.run(cmd, check=True)
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(numericalflowiteration_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, dp, i = C.c_void_p, C.c_size_t, C.POINTER(C.c_double), C.c_int
    for d in (1, 2, 3):
        f = getattr(L, f"nufi_b200_create_{d}d")
        f.argtypes = [vp, i, vp, i, C.POINTER(vp)]
        f.restype = i
    L.nufi_b200_destroy.argtypes = [vp]
    L.nufi_b200_destroy.restype = None
    L.nufi_b200_last_error.argtypes = [vp]
    L.nufi_b200_last_error.restype = C.c_char_p
    for name, args in {
        "compute_rho": [vp, sz, sz, sz], "download_rho": [vp, vp], "upload_phi": [vp, sz, vp],
        "compute_metrics": [vp, sz, sz, sz], "download_metrics": [vp, vp], "eval_rho_all": [vp, sz, vp],
        "solve_interpolate": [vp, sz, vp], "solve_interpolate_host": [vp, sz, vp, vp], "step": [vp, sz], "step_host": [vp, sz, vp, vp, vp, i],
        "download_energy": [vp, sz, sz, vp], "download_phi": [vp, sz, vp], "sync": [vp], "set_stream": [vp, vp],
        "rho_device": [vp, C.POINTER(vp)], "field_tail_device": [vp, sz, vp], "last_backtrace_ms": [vp, C.POINTER(C.c_float)],
        "set_variant": [vp, i], "measure_fp64_peak": [i, dp], "set_kernel_timing": [vp, i], "set_tile_nodes": [vp, i],
        "backtrace_time": [vp, dp, C.POINTER(C.c_uint64), i], "set_tail_variant": [vp, i],
        "download_history": [vp, sz, vp], "upload_history": [vp, sz, vp], "eval_f": [vp, sz, sz, vp, vp, i],
        "eval_field": [vp, sz, i, sz, vp, vp], "eval_phase_flow": [vp, sz, sz, vp, vp],
        "poisson_solve": [vp, vp, dp], "interpolate": [vp, vp, vp], "device_count": [C.POINTER(i)], "device_of": [vp],
        "group_create": [C.POINTER(vp), i, C.POINTER(vp)], "group_step": [vp, sz], "group_sync": [vp],
        "group_set_exchange": [vp, i],
        "peer_export": [vp, i, vp], "peer_attach": [vp, i, i, vp], "peer_step": [vp, sz], "peer_status": [vp, C.POINTER(i)],
        "peer_detach": [vp], "set_metrics_grid_1d": [vp, vp], "download_rho_full": [vp, vp],
    }.items():
        f = getattr(L, "nufi_b200_" + name)
        f.argtypes = args
        f.restype = i
    L.nufi_b200_launch_count.argtypes = [vp]
    L.nufi_b200_launch_count.restype = C.c_uint64
    L.nufi_b200_last_variant.argtypes = [vp]
    L.nufi_b200_last_variant.restype = C.c_char_p
    L.nufi_b200_last_tail_variant.argtypes = [vp]
    L.nufi_b200_last_tail_variant.restype = C.c_char_p
    L.nufi_b200_group_destroy.argtypes = [vp]
    L.nufi_b200_group_destroy.restype = None
    L.nufi_b200_group_last_error.argtypes = [vp]
    L.nufi_b200_group_last_error.restype = C.c_char_p
    L.nufi_b200_group_exchange.argtypes = [vp]
    L.nufi_b200_group_exchange.restype = C.c_char_p
    L.nufi_b200_version.restype = C.c_char_p
    _lib = L
    return L
