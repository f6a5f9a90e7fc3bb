"""numpy twin of the reference's ``TruckTrailerModel`` (python-files/truck_trailer_model.py).

The reference builds CasADi SX expressions; the CUDA solver has the kinematics and their derivatives
hand-coded (csrc/ttmpc_core.cuh), so this class only carries the constants and offers numeric versions of
the same methods for callers (plant simulation, plotting helpers).
"""
from __future__ import annotations

import numpy as np


class TruckTrailerModel:
    def __init__(self, params):
        self.num_state = 6  # q := (x, y, theta, psi, phi, v)     truck_trailer_model.py:4,12
        self.num_input = 2  # u := (a, omega)                     truck_trailer_model.py:5,14
        self._params = params

    @property
    def params(self):
        return self._params

    def f(self, q, u):
        """Continuous kinematics, truck_trailer_model.py:8-24."""
        L1, L2, M = self._params["L1"], self._params["L2"], self._params["M"]
        q = np.asarray(q, dtype=np.float64).reshape(-1)
        u = np.asarray(u, dtype=np.float64).reshape(-1)
        _, _, theta, psi, phi, v = q
        a, omega = u
        return np.array([
            v * np.cos(theta),
            v * np.sin(theta),
            v * np.tan(phi) / L1,
            -v * np.tan(phi) / L1 * (1 + M / L2 * np.cos(psi)) - v * np.sin(psi) / L2,
            omega,
            a,
        ])

    def compute_next_state(self, x_k, u_k):
        """Explicit Euler step, truck_trailer_model.py:26-29."""
        return np.asarray(x_k, dtype=np.float64).reshape(-1) + self.f(x_k, u_k) * self._params["dt"]

    # box H-representations and body centres (truck_trailer_model.py:31-72), used by the OBCA variant only
    def get_vehicle_Hrep(self):
        G = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])
        g = np.array([self._params["L1"] / 2, self._params["W1"] / 2, self._params["L1"] / 2, self._params["W1"] / 2])
        return G, g.reshape(4, 1)

    def get_trailer_Hrep(self):
        G = np.array([[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0], [0.0, -1.0]])
        g = np.array([self._params["L2"] / 2, self._params["W2"] / 2, self._params["L2"] / 2, self._params["W2"] / 2])
        return G, g.reshape(4, 1)

    def get_vehicle_center(self, x_rear, y_rear, heading):
        return (x_rear + np.cos(heading) * self._params["L1"] / 2, y_rear + np.sin(heading) * self._params["L1"] / 2)

    def get_trailer_center(self, x_rear, y_rear, heading, psi):
        xh = x_rear - np.cos(heading) * self._params["M"]
        yh = y_rear - np.sin(heading) * self._params["M"]
        return (xh - np.cos(heading + psi) * self._params["L2"] / 2, yh - np.sin(heading + psi) * self._params["L2"] / 2)
