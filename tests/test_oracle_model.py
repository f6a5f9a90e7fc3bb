"""Pins the oracle's model (a1/a2 of SURVEY.md section 8) to the reference's only golden data: the
planner output data/state_traj.txt + data/input_traj.txt is Euler-consistent with
truck_trailer_model.py:8-29 at dt=0.1 to 2.6e-13; derivatives are checked by finite differences."""
import os

import numpy as np

from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200 import tracking_preset
from oracle import oracle


def test_golden_trajectory_is_euler_consistent():
    cfg = tracking_preset(40, dt=0.1)
    S = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt")).T
    U = np.loadtxt(os.path.join(pb.DATA_DIR, "input_traj.txt")).T
    assert S.shape == (201, 6) and U.shape == (200, 2)
    worst = 0.0
    for k in range(200):
        f, _, _ = oracle.model(cfg, S[k], U[k], np.zeros(6))
        worst = max(worst, np.abs(S[k + 1] - S[k] - 0.1 * f).max())
    assert worst < 1e-12, worst
    # the host-side numpy model agrees with the oracle's
    d = pb.dynamics_defect(cfg, S, U)
    assert np.abs(d).max() < 1e-12


def test_planner_output_respects_bounds():
    cfg = tracking_preset(40)
    S = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt")).T
    U = np.loadtxt(os.path.join(pb.DATA_DIR, "input_traj.txt")).T
    assert (np.abs(S[:, 3]) <= np.pi / 3 + 1e-6).all()
    assert (np.abs(S[:, 4]) <= np.pi / 4 + 1e-6).all()
    assert (np.abs(U[:, 0]) <= 5 + 1e-6).all() and (np.abs(U[:, 1]) <= np.pi / 2 + 1e-6).all()
    assert cfg.x_ub[3] == np.pi / 3


def test_jacobian_and_hessian_vs_finite_differences():
    cfg = tracking_preset(40)
    rng = np.random.default_rng(1)
    for _ in range(20):
        q = rng.normal(0, 1, 6) * np.array([10, 10, 1.5, 0.5, 0.4, 4])
        u = rng.normal(0, 1, 2)
        lam = rng.normal(0, 1, 6)
        f0, Fx, H = oracle.model(cfg, q, u, lam)
        h = 1e-6
        Fx_fd = np.zeros((6, 6))
        H_fd = np.zeros((6, 6))
        for j in range(6):
            e = np.zeros(6); e[j] = h
            fp, Fp, _ = oracle.model(cfg, q + e, u, lam)
            fm, Fm, _ = oracle.model(cfg, q - e, u, lam)
            Fx_fd[:, j] = (fp - fm) / (2 * h)
            H_fd[:, j] = ((Fp - Fm) / (2 * h)).T @ lam
        assert np.abs(Fx - Fx_fd).max() < 1e-7
        assert np.abs(H - H_fd).max() < 1e-6
        assert np.abs(H - H.T).max() == 0.0
        assert f0[4] == u[1] and f0[5] == u[0]  # u = (a, omega): phi' = omega, v' = a
