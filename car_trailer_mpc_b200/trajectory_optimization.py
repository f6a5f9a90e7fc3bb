"""Drop-in for the reference's offline planner ``TrajectoryOptimization`` (python-files/trajectory_optimization.py).

Same constructor and ``plan(initial_state, goal_state) -> (states[6,N+1], inputs[2,N])`` (trajectory_optimization.py:11-29,
311-331; driven by trajectory_animation.py:41-110 with horizon 200, dt 0.1 and the 11 rectangles of obstacles.json).  The
CasADi/Ipopt call at :316-323 is replaced by one ``ttmpc_plan_batch`` call through the C ABI: the obstacle-aware NLP with
the goal-tracking cost of :175-183 (terminal weight 100 Q), the final-state box of :168-173 (+-1e-2) and the initial
trajectory of :227-274 -- cubic-spline interpolation of Hybrid-A* waypoints (positions, headings + pi/2, hitch angles;
steering angle, speed and inputs start at zero).  The reference reads the waypoints from ``initialize.json`` next to its
package; here they are an argument (``waypoints=``: that dictionary, or a path to such a file), and without them the guess
is the straight line of ``_generate_initial_trajectory_guess`` (:208-225).  One problem runs on one CTA of the GPU.
"""
from __future__ import annotations

import json

import numpy as np

from .config import Obstacles, planner_preset
from .mpc_control import ST_ACCEPTABLE, ST_CONVERGED, config_from_reference_args
from .solver import BatchSolver


def interpolate_waypoints(waypoints, num_output_nodes: int) -> np.ndarray:
    """interpolate_waypoints.py:5-27: cubic spline through evenly spaced waypoints, sampled evenly (two waypoints: a line)."""
    w = np.asarray(waypoints, dtype=np.float64)
    t_in = np.linspace(0.0, 1.0, len(w))
    t_out = np.linspace(0.0, 1.0, num_output_nodes)
    if len(w) < 3:
        return np.stack([np.interp(t_out, t_in, c) for c in w.reshape(len(w), -1).T], axis=-1).reshape((num_output_nodes,) + w.shape[1:])
    from scipy.interpolate import CubicSpline

    return CubicSpline(t_in, w)(t_out)


class TrajectoryOptimization:
    TERMINAL_WEIGHT = 100.0  # trajectory_optimization.py:181
    TERMINAL_BOX = 1e-2      # trajectory_optimization.py:170-171

    def __init__(self, dynamics, params, Q, R, state_bound, input_bound, obstacle_list, device: int = 0, waypoints=None,
                 geometric_start: bool = False):
        """``geometric_start`` (not in the reference, off by default): TTMPC_OBCA_GEOMETRIC_START, the OBCA duals start at the
        distance problems' multipliers for the poses of the initial trajectory (see mpc_control_obs.py of this package)."""
        self._dynamics = dynamics
        self._horizon = int(params["horizon"])
        for key in ("W1", "W2"):
            if key not in params:
                raise KeyError(f"params['{key}'] missing (trajectory_animation.py:49-53)")
        self._cfg = config_from_reference_args(dynamics, params, Q, R, state_bound, input_bound, planner_preset(self._horizon))
        self.obstacle_list = list(obstacle_list)
        self._obstacles = Obstacles.from_list(self.obstacle_list, W1=float(params["W1"]), W2=float(params["W2"]),
                                              geometric_start=geometric_start)
        self._solver = BatchSolver(self._cfg, device)
        self._waypoints = waypoints
        self.last_status = self.last_iterations = self.last_objective = None

    # ---- initial trajectories (states [N+1,6]; inputs start at zero) --------------------------------------------------
    def _generate_initial_trajectory_guess(self, initial_state, goal_state) -> np.ndarray:
        """trajectory_optimization.py:208-225: states on the straight line from start to goal, the last one at the goal"""
        N = self._horizon
        t = (np.arange(N + 1) / N)[:, None]
        g = (1.0 - t) * np.asarray(initial_state, float)[None] + t * np.asarray(goal_state, float)[None]
        g[N] = goal_state
        return g

    def _hybrid_a_star_initial_trajectory(self, waypoints) -> np.ndarray:
        """trajectory_optimization.py:227-274: spline through the planner's waypoints, headings shifted by pi/2 (:240),
        steering angle and speed at zero (:252-253); the reference samples `horizon` points and repeats the last one."""
        if isinstance(waypoints, str):
            with open(waypoints) as f:
                waypoints = json.load(f)
        N = self._horizon
        pos = interpolate_waypoints(waypoints["Positions"], N)
        hdg = interpolate_waypoints(np.asarray(waypoints["Headings"], float) + np.pi / 2.0, N)
        hit = interpolate_waypoints(waypoints["HitchAngles"], N)
        g = np.zeros((N + 1, 6))
        g[:N, 0:2], g[:N, 2], g[:N, 3] = pos, hdg, hit
        g[N] = g[N - 1]
        return g

    def plan(self, initial_state, goal_state, waypoints=None):
        N = self._horizon
        x0 = np.asarray(initial_state, dtype=np.float64).reshape(6)
        goal = np.asarray(goal_state, dtype=np.float64).reshape(6)
        wp = waypoints if waypoints is not None else self._waypoints
        guess = self._hybrid_a_star_initial_trajectory(wp) if wp is not None else self._generate_initial_trajectory_guess(x0, goal)
        z = np.zeros(8 * N + 6)
        for k in range(N + 1):
            z[8 * k:8 * k + 6] = guess[k]
        r = self._solver.plan(self._obstacles, x0[None], goal, self.TERMINAL_WEIGHT, self.TERMINAL_BOX, z[None])
        self.last_status, self.last_iterations = int(r["status"][0]), int(r["iters"][0])
        self.last_objective = float(r["obj"][0])
        zz = r["z"][0]
        states = np.stack([zz[8 * k:8 * k + 6] for k in range(N + 1)], axis=1)
        inputs = np.stack([zz[8 * k + 6:8 * k + 8] for k in range(N)], axis=1)
        return states, inputs  # (6, N+1), (2, N) like _split_decision_variables (:276-309)

    def converged(self) -> bool:
        return self.last_status in (ST_CONVERGED, ST_ACCEPTABLE)
