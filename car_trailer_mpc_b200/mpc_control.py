"""Drop-in for the reference's ``MPCTrackingControl`` (python-files/mpc_control.py).

Same constructor and ``solve`` signature (mpc_control.py:6-10,67-70); the CasADi/Ipopt call at
mpc_control.py:80-89 is replaced by one ``ttmpc_solve_batch`` call with B = 1 through the C ABI.
Failure behaviour is the reference's (mpc_control.py:106-110): print ``"Cannot find a solution!"`` and
return the last iterate.
"""
from __future__ import annotations

import numpy as np

from .config import ST_ACCEPTABLE, ST_CONVERGED, Config, tracking_preset
from .solver import BatchSolver


def _to_array(v) -> np.ndarray:
    """Accept numpy/list values and CasADi ``DM`` (anything with ``.full()``), as simulation.py:411-414 passes."""
    if hasattr(v, "full"):
        v = v.full()
    return np.asarray(v, dtype=np.float64).reshape(-1)


def config_from_reference_args(dynamics, params, Q, R, state_bound, input_bound, base: Config) -> Config:
    if getattr(dynamics, "num_state", None) != 6 or getattr(dynamics, "num_input", None) != 2:
        raise ValueError("dynamics must be the 6-state / 2-input truck-trailer model (truck_trailer_model.py:4-5)")
    for key in ("M", "L1", "L2", "dt", "horizon"):
        if key not in params:
            raise KeyError(f"params['{key}'] missing (simulation.py:391-395)")
    c = base.copy()
    c.horizon = int(params["horizon"])
    c.dt = float(params["dt"])
    c.L1, c.L2, c.M = float(params["L1"]), float(params["L2"]), float(params["M"])
    c.set_weights(np.asarray(Q, dtype=np.float64), np.asarray(R, dtype=np.float64))
    c.set_bounds(_to_array(state_bound["lb"]), _to_array(state_bound["ub"]),
                 _to_array(input_bound["lb"]), _to_array(input_bound["ub"]))
    return c


class MPCTrackingControl:
    def __init__(self, dynamics, params, Q, R, state_bound, input_bound, device: int = 0):
        self._dynamics = dynamics
        self._horizon = int(params["horizon"])
        self._num_state = 6
        self._num_input = 2
        self._cfg = config_from_reference_args(dynamics, params, Q, R, state_bound, input_bound,
                                               tracking_preset(self._horizon))
        self._solver = BatchSolver(self._cfg, device)
        self.last_status = None
        self.last_iterations = None
        self.last_objective = None

    def _success(self, status: int) -> bool:
        return status in (ST_CONVERGED, ST_ACCEPTABLE)

    def _solve_raw(self, initial_state, reference_states, reference_inputs, z_warm=None):
        N = self._horizon
        x0 = np.asarray(initial_state, dtype=np.float64).reshape(6)
        # the reference flattens stage-major via `.T.reshape((-1, 1))` (mpc_control.py:71-72)
        xs = np.ascontiguousarray(np.asarray(reference_states, dtype=np.float64).reshape(6, N + 1).T)
        us = np.ascontiguousarray(np.asarray(reference_inputs, dtype=np.float64).reshape(2, N).T)
        r = self._solver.solve(x0[None], xs[None], us[None], z_warm=None if z_warm is None else z_warm[None])
        self.last_status = int(r["status"][0])
        self.last_iterations = int(r["iters"][0])
        self.last_objective = float(r["obj"][0])
        return r["z"][0]

    def _split_decision_variables(self, z):
        """trajectory_planning.py:62-84: states [6,N+1], inputs [2,N] (column k = stage k), fresh arrays."""
        N = self._horizon
        body = z[: 8 * N].reshape(N, 8)
        states = np.empty((6, N + 1))
        states[:, :N] = body[:, :6].T
        states[:, N] = z[8 * N:]
        inputs = np.ascontiguousarray(body[:, 6:].T)
        return states, inputs

    def solve(self, initial_state, reference_states, reference_inputs):
        z = self._solve_raw(initial_state, reference_states, reference_inputs)
        if not self._success(self.last_status):
            print("Cannot find a solution!")
        return self._split_decision_variables(z)
