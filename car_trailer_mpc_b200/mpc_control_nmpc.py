"""Drop-in for the reference's ``TruckTrailerNMPC`` (python-files/mpc_control_nmpc.py).

Same NLP as ``MPCTrackingControl``; differences: loose tolerances
(mpc_control_nmpc.py:36-45), primal warm start from the shifted previous solution (:69-96), and
``(None, None)`` on failure without touching the stored warm start (:107-111).
"""
from __future__ import annotations

import numpy as np

from .config import nmpc_preset
from .mpc_control import MPCTrackingControl, config_from_reference_args
from .problem import shift_warm_start
from .solver import BatchSolver


class TruckTrailerNMPC(MPCTrackingControl):
    def __init__(self, dynamics, params, Q, R, state_bound, input_bound, device: int = 0,
                 shift_reference_bug: bool = True):
        self._dynamics = dynamics
        self._horizon = int(params["horizon"])
        self._num_state = 6
        self._num_input = 2
        self._cfg = config_from_reference_args(dynamics, params, Q, R, state_bound, input_bound,
                                               nmpc_preset(self._horizon))
        self._solver = BatchSolver(self._cfg, device)
        self._last_solution = None  # warm start between calls (mpc_control_nmpc.py:15)
        # True (default, as in MPCTrackingControlFuzzy -- both reference classes slice the same way) reproduces the
        # reference's mis-sliced warm-start tail (mpc_control_nmpc.py:83-87) bit for bit; False = the intended shift
        self._shift_reference_bug = bool(shift_reference_bug)
        self.last_status = None
        self.last_iterations = None
        self.last_objective = None

    def _shift_solution(self, vars_opt):
        return shift_warm_start(np.asarray(vars_opt, dtype=np.float64), self._horizon, self._shift_reference_bug)

    def solve(self, initial_state, reference_states, reference_inputs):
        guess = self._shift_solution(self._last_solution) if self._last_solution is not None else None
        z = self._solve_raw(initial_state, reference_states, reference_inputs, z_warm=guess)
        if not self._success(self.last_status):
            return None, None
        self._last_solution = z
        return self._split_decision_variables(z)
