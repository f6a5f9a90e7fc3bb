"""Loader for the TEST-ONLY host build of the device solver core (tools/kernel_emu.cpp)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libttmpc_emu.so")
_lib = None


_SO_SPEC = os.path.join(_HERE, "libttmpc_emu_spec.so")
_lib_spec = None


def _build(so, extra):
    src = os.path.join(_HERE, "kernel_emu.cpp")
    core = os.path.join(_HERE, "..", "car_trailer_mpc_b200", "csrc", "ttmpc_core.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-DTTMPC_BANK=64", *extra, "-fPIC", "-shared", "-o", so, src, "-lm"])
    return ctypes.CDLL(so)


def lib():
    """The core exactly as the shipped library compiles it (no experiment macros)."""
    global _lib
    if _lib is None:
        _lib = _build(_SO, [])
    return _lib


def lib_spec():
    """The experiment configuration -DTTMPC_SPECULATION=1 (speculative first line-search trial)."""
    global _lib_spec
    if _lib_spec is None:
        _lib_spec = _build(_SO_SPEC, ["-DTTMPC_SPECULATION=1"])
    return _lib_spec


def solve_batch(cfg, x_init, ref_states=None, ref_inputs=None, z_warm=None, k_index=None, traj_states=None,
                traj_inputs=None, force_generic=False, q_weights=None, r_weights=None):
    N = cfg.horizon
    x = np.ascontiguousarray(x_init, dtype=np.float64).reshape(-1, 6)
    B = x.shape[0]
    xs = us = ki = ts = tu = None
    T = 0
    if ref_states is not None:
        xs = np.ascontiguousarray(ref_states, dtype=np.float64).reshape(B, N + 1, 6)
        us = np.ascontiguousarray(ref_inputs, dtype=np.float64).reshape(B, N, 2)
    else:
        ki = np.ascontiguousarray(k_index, dtype=np.int32).reshape(B)
        ts = np.ascontiguousarray(traj_states, dtype=np.float64)
        tu = np.ascontiguousarray(traj_inputs, dtype=np.float64)
        T = tu.shape[0]
    qw = None if q_weights is None else np.ascontiguousarray(q_weights, dtype=np.float64).reshape(B, 6)
    rw = None if r_weights is None else np.ascontiguousarray(r_weights, dtype=np.float64).reshape(B, 2)
    zw = None if z_warm is None else np.ascontiguousarray(z_warm, dtype=np.float64).reshape(B, 8 * N + 6)
    z = np.empty((B, 8 * N + 6)); u0 = np.empty((B, 2)); obj = np.empty(B); kkt = np.empty((B, 3))
    it = np.empty(B, np.int32); st = np.empty(B, np.int32)
    dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int32)
    P = lambda a, t=dp: None if a is None else a.ctypes.data_as(t)
    L = lib_spec() if (int(force_generic) >> 3) & 3 else lib()   # bits 3-4: Params::speculate of the experiment build
    rc = L.ttmpc_emu_solve_batch(ctypes.byref(cfg), ctypes.c_int64(B), P(x), P(xs), P(us), P(ki, ip), P(ts), P(tu),
                                     ctypes.c_int32(T), P(zw), P(z), P(u0), P(obj), P(kkt), P(it, ip), P(st, ip),
                                     ctypes.c_int(int(force_generic)), P(qw), P(rw))
    if rc:
        raise RuntimeError(f"emu rc={rc}")
    return dict(z=z, u0=u0, obj=obj, kkt=kkt, iters=it, status=st)


_OSO = os.path.join(_HERE, "libttmpc_obca_emu.so")
_olib = None


def obca_lib():
    global _olib
    if _olib is None:
        src = os.path.join(_HERE, "obca_emu.cpp")
        deps = [src] + [os.path.join(_HERE, "..", "car_trailer_mpc_b200", "csrc", f) for f in ("ttmpc_core.cuh", "ttmpc_obca.cuh")]
        if not os.path.exists(_OSO) or os.path.getmtime(_OSO) < max(os.path.getmtime(f) for f in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-DTTMPC_BANK=64", "-fPIC", "-shared", "-o", _OSO, src, "-lm"])
        _olib = ctypes.CDLL(_OSO)
    return _olib


def plan_terminal_config(cfg, goal, terminal_weight, terminal_box):
    """configuration of the terminal stage of the planner's NLP, as ttmpc_plan_batch builds it"""
    import copy
    ct = copy.deepcopy(cfg)
    for i in range(36):
        ct.Q[i] = cfg.Q[i] * terminal_weight
    if terminal_box > 0:
        for i in range(6):
            ct.x_lb[i] = max(cfg.x_lb[i], goal[i] - terminal_box)
            ct.x_ub[i] = min(cfg.x_ub[i], goal[i] + terminal_box)
    return ct


def obca_plan_batch(cfg, obstacles, x_init, goal, terminal_weight=100.0, terminal_box=1e-2, z_guess=None, wide_warps=0):
    """host build of ttmpc_plan_batch (csrc/ttmpc.cu): same one-record shared trajectory, same terminal-stage parameters"""
    goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(6)
    x = np.ascontiguousarray(x_init, dtype=np.float64).reshape(-1, 6)
    ct = plan_terminal_config(cfg, goal, terminal_weight, terminal_box)
    zg = None if z_guess is None else np.ascontiguousarray(z_guess, dtype=np.float64).reshape(x.shape[0], 8 * cfg.horizon + 6)
    dp = ctypes.POINTER(ctypes.c_double)
    obca_lib().ttmpc_emu_obca_set_plan(ctypes.byref(ct), None if zg is None else zg.ctypes.data_as(dp))
    try:
        return obca_solve_batch(cfg, obstacles, x, k_index=np.zeros(x.shape[0], np.int32), traj_states=np.stack([goal, goal]),
                                traj_inputs=np.zeros((1, 2)), wide_warps=wide_warps)
    finally:
        obca_lib().ttmpc_emu_obca_set_plan(None, None)


def obca_solve_batch(cfg, obstacles, x_init, ref_states=None, ref_inputs=None, k_index=None, traj_states=None,
                     traj_inputs=None, wide_warps=0):
    """wide_warps > 0: the CTA-per-problem flavour of the kernel (stages dealt to that many virtual warps)."""
    obca_lib().ttmpc_emu_obca_set_wide(ctypes.c_int(int(wide_warps)))
    N = cfg.horizon
    x = np.ascontiguousarray(x_init, dtype=np.float64).reshape(-1, 6)
    B = x.shape[0]
    xs = us = ki = ts = tu = None
    T = 0
    if ref_states is not None:
        xs = np.ascontiguousarray(ref_states, dtype=np.float64).reshape(B, N + 1, 6)
        us = np.ascontiguousarray(ref_inputs, dtype=np.float64).reshape(B, N, 2)
    else:
        ki = np.ascontiguousarray(k_index, dtype=np.int32).reshape(B)
        ts = np.ascontiguousarray(traj_states, dtype=np.float64)
        tu = np.ascontiguousarray(traj_inputs, dtype=np.float64)
        T = tu.shape[0]
    z = np.empty((B, 8 * N + 6)); u0 = np.empty((B, 2)); obj = np.empty(B); kkt = np.empty((B, 3))
    it = np.empty(B, np.int32); st = np.empty(B, np.int32)
    dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int32)
    P = lambda a, t=dp: None if a is None else a.ctypes.data_as(t)
    rc = obca_lib().ttmpc_emu_obca_solve_batch(ctypes.byref(cfg), ctypes.byref(obstacles), ctypes.c_int64(B), P(x), P(xs),
                                               P(us), P(ki, ip), P(ts), P(tu), ctypes.c_int32(T), P(z), P(u0), P(obj),
                                               P(kkt), P(it, ip), P(st, ip))
    if rc:
        raise RuntimeError(f"obca emu rc={rc}")
    return dict(z=z, u0=u0, obj=obj, kkt=kkt, iters=it, status=st)


_TSO = os.path.join(_HERE, "libttmpc_team_emu.so")
_tlib = None


def team_lib():
    """Host build of the warp-cooperative solve (csrc/ttmpc_team.cuh): the 32 lanes of a warp run as fibers (tools/team_emu.cpp)."""
    global _tlib
    if _tlib is None:
        src = os.path.join(_HERE, "team_emu.cpp")
        deps = [src] + [os.path.join(_HERE, "..", "car_trailer_mpc_b200", "csrc", f) for f in ("ttmpc_core.cuh", "ttmpc_team.cuh")]
        if not os.path.exists(_TSO) or os.path.getmtime(_TSO) < max(os.path.getmtime(f) for f in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", _TSO, src, "-lm"])
        _tlib = ctypes.CDLL(_TSO)
    return _tlib


def team_solve_batch(cfg, x_init, ref_states=None, ref_inputs=None, z_warm=None, k_index=None, traj_states=None,
                     traj_inputs=None, lanes=16, flags=0):
    """lanes: lanes of the warp per problem (8, 16, 32); flags bit 0: generic-bounds variant, bit 1: dense-weights variant."""
    N = cfg.horizon
    x = np.ascontiguousarray(x_init, dtype=np.float64).reshape(-1, 6)
    B = x.shape[0]
    xs = us = ki = ts = tu = None
    T = 0
    if ref_states is not None:
        xs = np.ascontiguousarray(ref_states, dtype=np.float64).reshape(B, N + 1, 6)
        us = np.ascontiguousarray(ref_inputs, dtype=np.float64).reshape(B, N, 2)
    else:
        ki = np.ascontiguousarray(k_index, dtype=np.int32).reshape(B)
        ts = np.ascontiguousarray(traj_states, dtype=np.float64)
        tu = np.ascontiguousarray(traj_inputs, dtype=np.float64)
        T = tu.shape[0]
    zw = None if z_warm is None else np.ascontiguousarray(z_warm, dtype=np.float64).reshape(B, 8 * N + 6)
    z = np.empty((B, 8 * N + 6)); u0 = np.empty((B, 2)); obj = np.empty(B); kkt = np.empty((B, 3))
    it = np.empty(B, np.int32); st = np.empty(B, np.int32)
    dp = ctypes.POINTER(ctypes.c_double); ip = ctypes.POINTER(ctypes.c_int32)
    P = lambda a, t=dp: None if a is None else a.ctypes.data_as(t)
    rc = team_lib().ttmpc_team_emu_solve_batch(ctypes.byref(cfg), ctypes.c_int64(B), P(x), P(xs), P(us), P(ki, ip), P(ts),
                                               P(tu), ctypes.c_int32(T), P(zw), P(z), P(u0), P(obj), P(kkt), P(it, ip),
                                               P(st, ip), ctypes.c_int(int(lanes)), ctypes.c_int(int(flags)))
    if rc:
        raise RuntimeError(f"team emu rc={rc}")
    return dict(z=z, u0=u0, obj=obj, kkt=kkt, iters=it, status=st)
