"""ctypes binding of ``libttmpc.so`` (the C ABI declared in include/ttmpc.h).

The product has exactly one compute path: the CUDA library.  If it is missing or no CUDA device is
usable this module raises -- there is no CPU fallback and nothing here imports ``oracle/``.
"""
from __future__ import annotations

import ctypes
import os

from .config import Config

_LIB_PATH = os.environ.get("TTMPC_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libttmpc.so")
_lib = None

c_dp = ctypes.c_void_p  # raw device/host pointers are passed as integers

E_NODEV = -19


class TTMPCError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise TTMPCError(
            f"{_LIB_PATH} not found: build it with `python -m car_trailer_mpc_b200.build` "
            "(there is no CPU fallback for the solve)"
        )
    L = ctypes.CDLL(_LIB_PATH)
    H = ctypes.c_void_p
    L.ttmpc_version.restype = ctypes.c_char_p
    L.ttmpc_default_config.argtypes = [ctypes.POINTER(Config), ctypes.c_int32]
    L.ttmpc_default_config.restype = None
    L.ttmpc_create.argtypes = [ctypes.POINTER(Config), ctypes.c_int, ctypes.POINTER(H)]
    L.ttmpc_create.restype = ctypes.c_int
    L.ttmpc_destroy.argtypes = [H]
    L.ttmpc_destroy.restype = ctypes.c_int
    L.ttmpc_last_error.argtypes = [H]
    L.ttmpc_last_error.restype = ctypes.c_char_p
    L.ttmpc_solve_batch.argtypes = [H, ctypes.c_int64] + [c_dp] * 10 + [ctypes.c_void_p]
    L.ttmpc_solve_batch.restype = ctypes.c_int
    L.ttmpc_solve_batch_weighted.argtypes = [H, ctypes.c_int64] + [c_dp] * 12 + [ctypes.c_void_p]
    L.ttmpc_solve_batch_weighted.restype = ctypes.c_int
    L.ttmpc_solve_batch_shared.argtypes = (
        [H, ctypes.c_int64, c_dp, c_dp, c_dp, c_dp, ctypes.c_int32] + [c_dp] * 7 + [ctypes.c_void_p]
    )
    L.ttmpc_solve_batch_shared.restype = ctypes.c_int
    L.ttmpc_solve_batch_multi.argtypes = (
        [H, ctypes.c_int64, c_dp, c_dp, c_dp, c_dp, c_dp, ctypes.c_int32, ctypes.c_int32] + [c_dp] * 7 + [ctypes.c_void_p]
    )
    L.ttmpc_solve_batch_multi.restype = ctypes.c_int
    L.ttmpc_sync.argtypes = [H]
    L.ttmpc_sync.restype = ctypes.c_int
    L.ttmpc_host_pipeline_ms.argtypes = [H]
    L.ttmpc_host_pipeline_ms.restype = ctypes.c_double
    L.ttmpc_obca_solve_batch.argtypes = [H, ctypes.c_void_p, ctypes.c_int64] + [c_dp] * 9 + [ctypes.c_void_p]
    L.ttmpc_obca_solve_batch.restype = ctypes.c_int
    L.ttmpc_obca_solve_batch_shared.argtypes = (
        [H, ctypes.c_void_p, ctypes.c_int64, c_dp, c_dp, c_dp, c_dp, ctypes.c_int32] + [c_dp] * 6 + [ctypes.c_void_p]
    )
    L.ttmpc_obca_solve_batch_shared.restype = ctypes.c_int
    L.ttmpc_plan_batch.argtypes = ([H, ctypes.c_void_p, ctypes.c_int64, c_dp, ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
                                   + [c_dp] * 7 + [ctypes.c_void_p])
    L.ttmpc_plan_batch.restype = ctypes.c_int
    L.ttmpc_shift_warm_start.argtypes = [H, ctypes.c_int64, c_dp, c_dp, ctypes.c_int32, ctypes.c_void_p]
    L.ttmpc_shift_warm_start.restype = ctypes.c_int
    L.ttmpc_plant_step.argtypes = [H, ctypes.c_int64, c_dp, c_dp, c_dp, c_dp, ctypes.c_double, c_dp, ctypes.c_void_p]
    L.ttmpc_plant_step.restype = ctypes.c_int
    L.ttmpc_episode_batch.argtypes = [H, ctypes.c_int64, c_dp, c_dp, c_dp, c_dp, ctypes.c_int32, c_dp, ctypes.c_int32, c_dp,
                                      ctypes.c_int32, ctypes.c_uint64, c_dp, c_dp, ctypes.c_void_p]
    L.ttmpc_episode_batch.restype = ctypes.c_int
    L.ttmpc_last_solve_lanes.argtypes = [H]
    L.ttmpc_last_solve_lanes.restype = ctypes.c_int32
    L.ttmpc_launch_count.argtypes = [H]
    L.ttmpc_launch_count.restype = ctypes.c_int64
    L.ttmpc_kernel_name.argtypes = [H, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]
    L.ttmpc_kernel_name.restype = ctypes.c_char_p
    L.ttmpc_measure_fp64_peak.argtypes = [H, ctypes.c_void_p]
    L.ttmpc_measure_fp64_peak.restype = ctypes.c_double
    _lib = L
    return L


EXPORTS = [
    "ttmpc_default_config", "ttmpc_create", "ttmpc_destroy", "ttmpc_last_error", "ttmpc_version",
    "ttmpc_solve_batch", "ttmpc_solve_batch_weighted", "ttmpc_solve_batch_shared",
    "ttmpc_obca_solve_batch", "ttmpc_obca_solve_batch_shared", "ttmpc_shift_warm_start", "ttmpc_plant_step", "ttmpc_episode_batch",
    "ttmpc_launch_count", "ttmpc_kernel_name", "ttmpc_measure_fp64_peak", "ttmpc_last_solve_lanes", "ttmpc_solve_batch_multi", "ttmpc_sync", "ttmpc_host_pipeline_ms", "ttmpc_plan_batch",
]
