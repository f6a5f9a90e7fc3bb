"""ctypes front end of the CPU oracle ``libttmpc_oracle.so`` -- TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
bench.py import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from car_trailer_mpc_b200.config import Config

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libttmpc_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ttmpc_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ttmpc.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libttmpc_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.ttmpc_oracle_solve_batch.restype = ctypes.c_int
        _lib.ttmpc_oracle_solve.restype = ctypes.c_int
    return _lib


def _p(a, t=_dp):
    return None if a is None else a.ctypes.data_as(t)


def solve_batch(cfg: Config, x_init, ref_states, ref_inputs, z_warm=None, nthreads: int = 1):
    """Solve B problems on the CPU. Arrays stage-major: x_init [B,6], ref_states [B,N+1,6], ref_inputs [B,N,2].

    Returns dict(z [B,8N+6], u0 [B,2], obj [B], kkt [B,3], iters [B], status [B]).
    """
    N = cfg.horizon
    x_init = np.ascontiguousarray(x_init, dtype=np.float64).reshape(-1, 6)
    B = x_init.shape[0]
    ref_states = np.ascontiguousarray(ref_states, dtype=np.float64).reshape(B, N + 1, 6)
    ref_inputs = np.ascontiguousarray(ref_inputs, dtype=np.float64).reshape(B, N, 2)
    if z_warm is not None:
        z_warm = np.ascontiguousarray(z_warm, dtype=np.float64).reshape(B, 8 * N + 6)
    z = np.empty((B, 8 * N + 6))
    u0 = np.empty((B, 2))
    obj = np.empty(B)
    kkt = np.empty((B, 3))
    iters = np.empty(B, dtype=np.int32)
    status = np.empty(B, dtype=np.int32)
    rc = lib().ttmpc_oracle_solve_batch(
        ctypes.byref(cfg), ctypes.c_int64(B), _p(x_init), _p(ref_states), _p(ref_inputs), _p(z_warm),
        _p(z), _p(u0), _p(obj), _p(kkt), _p(iters, _ip), _p(status, _ip), ctypes.c_int(nthreads),
    )
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return dict(z=z, u0=u0, obj=obj, kkt=kkt, iters=iters, status=status)


def solve(cfg: Config, x_init, ref_states, ref_inputs, z_warm=None):
    r = solve_batch(cfg, np.asarray(x_init)[None], np.asarray(ref_states)[None], np.asarray(ref_inputs)[None],
                    None if z_warm is None else np.asarray(z_warm)[None])
    return {k: v[0] for k, v in r.items()}


def model(cfg: Config, q, u, lam):
    """f(q,u) [6], df/dq [6,6], sum_i lam_i d2f_i/dq2 [6,6] -- for the derivative tests."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    lam = np.ascontiguousarray(lam, dtype=np.float64)
    f = np.empty(6)
    Fx = np.empty((6, 6))
    H = np.empty((6, 6))
    lib().ttmpc_oracle_model(ctypes.byref(cfg), _p(q), _p(u), _p(lam), _p(f), _p(Fx), _p(H))
    return f, Fx, H


def window(S, U, k: int, N: int):
    S = np.ascontiguousarray(S, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    xs = np.empty((N + 1, 6))
    us = np.empty((N, 2))
    lib().ttmpc_oracle_window(_p(S), _p(U), ctypes.c_int(U.shape[0]), ctypes.c_int(k), ctypes.c_int(N), _p(xs), _p(us))
    return xs, us


def shift(z, N: int, mode: int):
    z = np.ascontiguousarray(z, dtype=np.float64)
    out = np.empty_like(z)
    lib().ttmpc_oracle_shift(_p(z), ctypes.c_int(N), ctypes.c_int(mode), _p(out))
    return out


def plant_step(cfg: Config, q, u, disturb=None, noise=None, noise_scale: float = 0.0):
    q = np.ascontiguousarray(q, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    d = None if disturb is None else np.ascontiguousarray(disturb, dtype=np.float64)
    n = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
    out = np.empty(6)
    lib().ttmpc_oracle_plant_step(ctypes.byref(cfg), _p(q), _p(u), _p(d), _p(n), ctypes.c_double(noise_scale), _p(out))
    return out
