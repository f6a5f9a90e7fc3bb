"""The controller switch of the reference's closed loop (python-files/simulation.py:416-436, 501-512, 524-526) as an
object with the controllers' own ``solve`` signature.

``simulation.py`` builds two controllers (obstacle-aware / plain; with ``USE_OBS_MPC`` both are obstacle-aware, otherwise
both plain) and, before every solve, tests the PREVIOUS solve's predicted states -- the reference window on the first
step -- against the obstacle list: any colliding stage selects the obstacle-aware controller for this step, which then
prints ``"Using obstacle-aware MPC"``.  The states returned by the chosen controller become the next step's test
trajectory.  Wrapping that in ``solve`` lets the headless drivers (``closed_loop.simulate_single``) run the reference's
switching loop with any pair of controllers, e.g. the ``MPCTrackingControlObs`` / ``MPCTrackingControl`` shims.
"""
from __future__ import annotations

import numpy as np

from .collision import check_trajectory_collision


class SwitchingController:
    def __init__(self, controller_obs, controller_no_obs, params: dict, obstacle_list, verbose: bool = True):
        self.controller_obs = controller_obs
        self.controller_no_obs = controller_no_obs
        self.params = params
        self.obstacle_list = list(obstacle_list)
        self.verbose = verbose
        self.prev_mpc_prediction = None      # simulation.py:470 / :526
        self.used_obstacle_aware = []        # one flag per solve (the reference only prints)

    def reset(self):
        self.prev_mpc_prediction = None
        self.used_obstacle_aware = []

    def solve(self, initial_state, reference_states, reference_inputs):
        check = self.prev_mpc_prediction if self.prev_mpc_prediction is not None else np.asarray(reference_states)
        needs = check_trajectory_collision(check, self.params, self.obstacle_list)
        controller = self.controller_obs if needs else self.controller_no_obs
        if needs and self.verbose:
            print("Using obstacle-aware MPC")
        self.used_obstacle_aware.append(bool(needs))
        states, inputs = controller.solve(initial_state, reference_states, reference_inputs)
        if states is not None:               # the reference stores np.array(states) unconditionally (:526); its
            self.prev_mpc_prediction = np.array(states)   # controllers on this path never return None
        return states, inputs
