"""Solver configuration: the ctypes mirror of ``ttmpc_config`` (include/ttmpc.h) and the presets
of the reference's two controllers.

Numbers follow the reference drivers:
  * ``MPCTrackingControl`` preset  -- python-files/simulation.py:388-414 (Q=I6, R=10*I2, bounds) with
    Ipopt defaults + ``max_iter 5000`` (python-files/mpc_control.py:35-39).
  * ``TruckTrailerNMPC`` preset    -- python-files/simulation_nmpc.py:124-148 (Q=diag(1,1,2,3,1,1),
    R=diag(5,8), v in [-8,8], a in [-4,4]) with ``tol 1e-3, acceptable_tol 1e-2, acceptable_iter 5,
    max_iter 2000`` (python-files/mpc_control_nmpc.py:36-45).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

NX = 6
NU = 2
MAX_HORIZON = 256
MAX_OBSTACLES = 16

# status codes (include/ttmpc.h)
ST_CONVERGED, ST_ACCEPTABLE, ST_MAX_ITER, ST_LINESEARCH, ST_NUMERIC, ST_INFEASIBLE_X0 = range(6)
STATUS_NAMES = {
    ST_CONVERGED: "converged",
    ST_ACCEPTABLE: "acceptable",
    ST_MAX_ITER: "max_iter",
    ST_LINESEARCH: "linesearch",
    ST_NUMERIC: "numeric",
    ST_INFEASIBLE_X0: "infeasible_x0",
}

FLAG_HOST_POINTERS = 0x1
FLAG_SYNC = 0x2
FLAG_ASYNC_HOST = 0x8


class Config(ctypes.Structure):
    """Plain-old-data twin of ``struct ttmpc_config``; field order and types must match the header."""

    _fields_ = [
        ("horizon", ctypes.c_int32),
        ("max_iter", ctypes.c_int32),
        ("acceptable_iter", ctypes.c_int32),
        ("flags", ctypes.c_uint32),
        ("dt", ctypes.c_double),
        ("L1", ctypes.c_double),
        ("L2", ctypes.c_double),
        ("M", ctypes.c_double),
        ("Q", ctypes.c_double * 36),
        ("R", ctypes.c_double * 4),
        ("x_lb", ctypes.c_double * 6),
        ("x_ub", ctypes.c_double * 6),
        ("u_lb", ctypes.c_double * 2),
        ("u_ub", ctypes.c_double * 2),
        ("tol", ctypes.c_double),
        ("acceptable_tol", ctypes.c_double),
        ("mu_init", ctypes.c_double),
    ]

    # -- helpers -------------------------------------------------------------------------------
    def set_weights(self, Q, R) -> None:
        Q = np.asarray(Q, dtype=np.float64).reshape(6, 6)
        R = np.asarray(R, dtype=np.float64).reshape(2, 2)
        self.Q[:] = Q.ravel().tolist()
        self.R[:] = R.ravel().tolist()

    def set_bounds(self, x_lb, x_ub, u_lb, u_ub) -> None:
        self.x_lb[:] = np.asarray(x_lb, dtype=np.float64).ravel().tolist()
        self.x_ub[:] = np.asarray(x_ub, dtype=np.float64).ravel().tolist()
        self.u_lb[:] = np.asarray(u_lb, dtype=np.float64).ravel().tolist()
        self.u_ub[:] = np.asarray(u_ub, dtype=np.float64).ravel().tolist()

    @property
    def nz(self) -> int:
        """Length of the reference's decision vector, 8N+6 (trajectory_planning.py:38-60)."""
        return 8 * self.horizon + 6

    def Qm(self) -> np.ndarray:
        return np.array(self.Q[:], dtype=np.float64).reshape(6, 6)

    def Rm(self) -> np.ndarray:
        return np.array(self.R[:], dtype=np.float64).reshape(2, 2)

    def copy(self) -> "Config":
        c = Config()
        ctypes.memmove(ctypes.byref(c), ctypes.byref(self), ctypes.sizeof(Config))
        return c


def tracking_preset(horizon: int = 40, dt: float = 0.05, max_iter: int = 5000) -> Config:
    """``MPCTrackingControl`` as configured by simulation.py:388-414 (Ipopt defaults, tol 1e-8)."""
    c = Config()
    c.horizon = horizon
    c.max_iter = max_iter
    c.acceptable_iter = 15
    c.flags = 0
    c.dt = dt
    c.L1, c.L2, c.M = 7.05, 12.45, 0.15
    c.set_weights(np.eye(6), 10.0 * np.eye(2))
    inf = math.inf
    c.set_bounds(
        [-inf, -inf, -math.pi, -math.pi / 3.0, -math.pi / 4.0, -10.0],
        [inf, inf, math.pi, math.pi / 3.0, math.pi / 4.0, 10.0],
        [-5.0, -math.pi / 2.0],
        [5.0, math.pi / 2.0],
    )
    c.tol = 1e-8
    c.acceptable_tol = 1e-6
    c.mu_init = 0.1
    return c


def planner_preset(horizon: int = 200, dt: float = 0.1, max_iter: int = 5000) -> Config:
    """``TrajectoryOptimization`` as configured by trajectory_animation.py:41-79 (Q = I, R = 10 I, heading unbounded,
    speed in [-5, 10]) and trajectory_optimization.py:196-199 (Ipopt defaults, max_iter 5000)."""
    c = tracking_preset(horizon, dt, max_iter)
    inf = math.inf
    c.set_bounds(
        [-inf, -inf, -inf, -math.pi / 3.0, -math.pi / 4.0, -5.0],
        [inf, inf, inf, math.pi / 3.0, math.pi / 4.0, 10.0],
        [-5.0, -math.pi / 2.0],
        [5.0, math.pi / 2.0],
    )
    return c


def nmpc_preset(horizon: int = 30, dt: float = 0.05, max_iter: int = 2000) -> Config:
    """``TruckTrailerNMPC`` as configured by simulation_nmpc.py:124-148 + mpc_control_nmpc.py:36-45."""
    c = tracking_preset(horizon, dt, max_iter)
    c.acceptable_iter = 5
    c.set_weights(np.diag([1.0, 1.0, 2.0, 3.0, 1.0, 1.0]), np.diag([5.0, 8.0]))
    inf = math.inf
    c.set_bounds(
        [-inf, -inf, -math.pi, -math.pi / 3.0, -math.pi / 4.0, -8.0],
        [inf, inf, math.pi, math.pi / 3.0, math.pi / 4.0, 8.0],
        [-4.0, -math.pi / 2.0],
        [4.0, math.pi / 2.0],
    )
    c.tol = 1e-3
    c.acceptable_tol = 1e-2
    return c


OBCA_NO_RECOVERY = 1  # TTMPC_OBCA_NO_RECOVERY
OBCA_GEOMETRIC_START = 2  # TTMPC_OBCA_GEOMETRIC_START (opt-in: not the reference's starting point for the OBCA duals)


class Obstacles(ctypes.Structure):
    """Plain-old-data twin of ``struct ttmpc_obstacles``: the obstacle list of ``MPCTrackingControlObs``
    (python-files/mpc_control_obs.py:8-30) as ``{centre x, centre y, width, height}`` rows, the body widths
    ``params['W1'], params['W2']`` (simulation.py:393) and the safety distance (mpc_control_obs.py:67)."""

    _fields_ = [
        ("count", ctypes.c_int32),
        ("flags", ctypes.c_int32),  # 0 or OBCA_NO_RECOVERY
        ("rect", (ctypes.c_double * 4) * MAX_OBSTACLES),
        ("W1", ctypes.c_double),
        ("W2", ctypes.c_double),
        ("d_min", ctypes.c_double),
    ]

    @classmethod
    def from_list(cls, obstacle_list, W1: float = 3.05, W2: float = 2.95, d_min: float = 0.2,
                  recover: bool = True, geometric_start: bool = False) -> "Obstacles":
        """``obstacle_list``: dicts with ``center`` / ``width`` / ``height`` (get_obstacles.py:5-33) or 4-tuples."""
        if not 1 <= len(obstacle_list) <= MAX_OBSTACLES:
            raise ValueError(f"between 1 and {MAX_OBSTACLES} obstacles are supported")
        o = cls()
        o.count = len(obstacle_list)
        for i, ob in enumerate(obstacle_list):
            if isinstance(ob, dict):
                row = (ob["center"][0], ob["center"][1], ob["width"], ob["height"])
            else:
                row = tuple(ob)
            for j in range(4):
                o.rect[i][j] = float(row[j])
        o.W1, o.W2, o.d_min = float(W1), float(W2), float(d_min)
        o.flags = (0 if recover else OBCA_NO_RECOVERY) | (OBCA_GEOMETRIC_START if geometric_start else 0)
        return o

    def as_list(self):
        return [dict(center=(self.rect[i][0], self.rect[i][1]), width=self.rect[i][2], height=self.rect[i][3])
                for i in range(self.count)]


def parking_lot_obstacles():
    """The 11 rectangles of python-files/obstacles.json as get_obstacles.py:5-33 returns them."""
    rects = [(-15.0, 10.0, 30.0, 20.0), (75.0, 10.0, 30.0, 20.0)]
    rects += [(c, 10.0, 5.0, 20.0) for c in (3.5, 9.5, 21.5, 27.5, 33.5, 39.5, 45.5, 51.5, 57.5)]
    return [dict(center=(r[0], r[1]), width=r[2], height=r[3]) for r in rects]
