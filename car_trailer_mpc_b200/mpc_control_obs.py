"""Drop-in for the reference's ``MPCTrackingControlObs`` (python-files/mpc_control_obs.py), the obstacle-aware
(OBCA) controller.

Same constructor and ``solve`` signature (mpc_control_obs.py:8-30,282-285); the CasADi/Ipopt call at
mpc_control_obs.py:296-305 is replaced by one ``ttmpc_obca_solve_batch`` call with B = 1 through the C ABI.  As in the
reference every solve is a cold start at the reference window with fixed dual guesses (:216-239), ``_last_solution``
is kept (:307), a failed solve prints ``"Cannot find a solution!"`` and the last iterate is returned (:319-322).  With
an empty obstacle list the reference builds the plain tracking NLP (:181-188); so does this class.

One keyword the reference does not have: ``geometric_start=True`` (``TTMPC_OBCA_GEOMETRIC_START``) starts the OBCA duals at
the multipliers of the distance problems for the pose of the reference window instead of the reference's constants
(:226-237).  Same NLP, same tolerances; a third to a fifth of the iterations (a solve of about 1.5 ms instead of 9 ms);
off by default because it is not the reference's iterate path (DESIGN.md section 3b has the measurements, including the
5 % of passage windows where the two starts end in different local solutions).
"""
from __future__ import annotations

import numpy as np

from .config import Obstacles
from .mpc_control import MPCTrackingControl


class MPCTrackingControlObs(MPCTrackingControl):
    def __init__(self, dynamics, params, Q, R, state_bound, input_bound, obstacle_list=None, device: int = 0,
                 geometric_start: bool = False):
        super().__init__(dynamics, params, Q, R, state_bound, input_bound, device=device)
        self.obstacle_list = list(obstacle_list) if obstacle_list is not None else []
        self._obstacles = None
        if self.obstacle_list:
            for key in ("W1", "W2"):
                if key not in params:
                    raise KeyError(f"params['{key}'] missing (simulation.py:393)")
            self._obstacles = Obstacles.from_list(self.obstacle_list, W1=float(params["W1"]), W2=float(params["W2"]),
                                                  geometric_start=geometric_start)
        self._last_solution = None

    def _solve_raw(self, initial_state, reference_states, reference_inputs, z_warm=None):
        if self._obstacles is None:
            return super()._solve_raw(initial_state, reference_states, reference_inputs)
        N = self._horizon
        x0 = np.asarray(initial_state, dtype=np.float64).reshape(6)
        xs = np.ascontiguousarray(np.asarray(reference_states, dtype=np.float64).reshape(6, N + 1).T)
        us = np.ascontiguousarray(np.asarray(reference_inputs, dtype=np.float64).reshape(2, N).T)
        r = self._solver.solve_obca(self._obstacles, x0[None], xs[None], us[None])
        self.last_status = int(r["status"][0])
        self.last_iterations = int(r["iters"][0])
        self.last_objective = float(r["obj"][0])
        return r["z"][0]

    def solve(self, initial_state, reference_states, reference_inputs):
        z = self._solve_raw(initial_state, reference_states, reference_inputs)
        self._last_solution = z
        if not self._success(self.last_status):
            print("Cannot find a solution!")
        return self._split_decision_variables(z)
