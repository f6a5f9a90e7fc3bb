"""Drop-in for the reference's ``MPCTrackingControlFuzzy`` (python-files/mpc_control_fuzzy.py).

Same NLP as the other controllers with per-solve weight scalings ``Q_w = diag(q) Q diag(q)``,
``R_w = diag(r) R diag(r)`` passed as parameters (mpc_control_fuzzy.py:21-31,51-60) -- here the per-problem
weight inputs of ``ttmpc_solve_batch_weighted``.  The rule base (:90-119), the shifted warm start (:72-88, the
same slicing as the NMPC controller), the retry with unit weights (:145-161) and the ``(None, None)`` failure return
are the reference's.
"""
from __future__ import annotations

import numpy as np

from .config import nmpc_preset
from .mpc_control import MPCTrackingControl, config_from_reference_args
from .problem import shift_warm_start
from .solver import BatchSolver


def fuzzy_weights(current_state, reference_states):
    """``_compute_fuzzy_weights`` (mpc_control_fuzzy.py:90-119): scale Q/R with hitch angle and reversing."""
    current_state = np.asarray(current_state, dtype=np.float64).reshape(-1)
    reference_states = np.asarray(reference_states, dtype=np.float64)
    psi = float(current_state[3])
    v = float(current_state[5])
    ref_v = float(reference_states[5, 0]) if reference_states.size > 0 else 0.0
    hitch_soft = 0.35
    hitch_norm = min(abs(psi) / hitch_soft, 1.0)
    reversing = (ref_v < -0.1) or (v < -0.1)
    q = np.ones(6)
    r = np.ones(2)
    hitch_gain = 1.0 + 2.0 * hitch_norm
    steer_gain = 1.0 + 1.2 * hitch_norm
    steer_rate_gain = 1.0 + 1.5 * hitch_norm
    if reversing:
        hitch_gain *= 1.1
        steer_gain *= 1.1
        steer_rate_gain *= 1.2
    q[2] = max(1.0, steer_gain)  # theta
    q[3] = max(1.0, hitch_gain)  # psi (hitch)
    q[4] = max(1.0, steer_gain)  # phi
    r[1] = max(1.0, steer_rate_gain)  # steering rate input
    return np.clip(q, 1.0, 3.5), np.clip(r, 1.0, 3.5)


class MPCTrackingControlFuzzy(MPCTrackingControl):
    def __init__(self, dynamics, params, Q, R, state_bound, input_bound, device: int = 0,
                 shift_reference_bug: bool = True):
        self._dynamics = dynamics
        self._horizon = int(params["horizon"])
        self._num_state = 6
        self._num_input = 2
        # same Ipopt options as the NMPC controller (mpc_control_fuzzy.py:41-50)
        self._cfg = config_from_reference_args(dynamics, params, Q, R, state_bound, input_bound, nmpc_preset(self._horizon))
        self._solver = BatchSolver(self._cfg, device)
        self._last_solution = None
        self._shift_reference_bug = bool(shift_reference_bug)
        self.last_status = self.last_iterations = self.last_objective = None
        self.last_weights = None

    def _compute_fuzzy_weights(self, current_state, reference_states):
        return fuzzy_weights(current_state, reference_states)

    def _solve_weighted(self, initial_state, reference_states, reference_inputs, guess, q, r):
        N = self._horizon
        x0 = np.asarray(initial_state, dtype=np.float64).reshape(1, 6)
        xs = np.ascontiguousarray(np.asarray(reference_states, dtype=np.float64).reshape(6, N + 1).T)[None]
        us = np.ascontiguousarray(np.asarray(reference_inputs, dtype=np.float64).reshape(2, N).T)[None]
        out = self._solver.solve(x0, xs, us, z_warm=None if guess is None else guess[None],
                                 q_weights=np.asarray(q, dtype=np.float64)[None], r_weights=np.asarray(r, dtype=np.float64)[None])
        self.last_status = int(out["status"][0])
        self.last_iterations = int(out["iters"][0])
        self.last_objective = float(out["obj"][0])
        return out["z"][0]

    def solve(self, initial_state, reference_states, reference_inputs):
        guess = None
        if self._last_solution is not None:
            guess = shift_warm_start(self._last_solution, self._horizon, self._shift_reference_bug)
        q, r = self._compute_fuzzy_weights(initial_state, reference_states)
        self.last_weights = (q, r)
        z = self._solve_weighted(initial_state, reference_states, reference_inputs, guess, q, r)
        if not self._success(self.last_status):  # retry with unit weights (mpc_control_fuzzy.py:145-161)
            z = self._solve_weighted(initial_state, reference_states, reference_inputs, guess, np.ones(6), np.ones(2))
        if not self._success(self.last_status):
            return None, None
        self._last_solution = z
        return self._split_decision_variables(z)
