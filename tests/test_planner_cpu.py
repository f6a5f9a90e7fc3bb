"""Offline planner (trajectory_optimization.py as a long-horizon single-problem solve), CPU side: the
kernel core compiled for the host (per-pair condensation + Riccati, terminal stage with its own bounds and weight)
against the oracle's generic block-tridiagonal LDL' on the planner's NLP, up to the reference's own size (horizon 200,
11 rectangles, 37 k variables), and solver-independent properties of the result."""
import os
import sys

import numpy as np
import pytest

import geometry
from obca_common import Z_TOL, golden_cases, split_z
from parity import OBJ_REL_TOL, U0_ABS_TOL

from car_trailer_mpc_b200 import planner_preset
from car_trailer_mpc_b200.config import Obstacles
from oracle import obca_oracle as ob

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emu  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "planner_cases.npz")
CASES = golden_cases(GOLD)
TERM_W, TERM_BOX = 100.0, 1e-2


def guess_z(c):
    N = int(c["horizon"])
    z = np.zeros(8 * N + 6)
    for k in range(N + 1):
        z[8 * k:8 * k + 6] = c["guess"][k]
    return z


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_kernel_core_matches_oracle_on_the_planner_nlp(c):
    N = int(c["horizon"])
    cfg = planner_preset(N)
    cfg.max_iter = 1000
    obs = Obstacles.from_list([tuple(r) for r in c["rects"]])
    r = emu.obca_plan_batch(cfg, obs, c["x_init"][None], c["goal"], TERM_W, TERM_BOX, guess_z(c)[None], wide_warps=8 if N >= 100 else 0)
    assert r["status"][0] == 0 and int(c["status"]) == 0
    xs, us = split_z(r["z"][0], N)
    assert np.abs(us[0] - c["inputs"][0]).max() <= U0_ABS_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    # solver-independent properties: dynamics, final-state box, bounds, true clearance
    from car_trailer_mpc_b200 import problem as pb
    assert np.abs(pb.dynamics_defect(cfg, xs[None], us[None])).max() <= 1e-8
    assert np.array_equal(xs[0], c["x_init"])
    # Ipopt relaxes every bound by 1e-8 * max(1, |bound|)
    assert (np.abs(xs[-1] - c["goal"]) <= TERM_BOX + 1.01e-8 * np.maximum(1.0, np.abs(c["goal"]) + TERM_BOX)).all()
    assert (np.abs(xs[1:, 3]) <= np.pi / 3 + 2e-8).all() and (np.abs(xs[1:, 4]) <= np.pi / 4 + 2e-8).all()
    assert (xs[1:, 5] >= -5 - 6e-8).all() and (xs[1:, 5] <= 10 + 1.1e-7).all()
    assert geometry.clearance(xs, [tuple(q) for q in c["rects"]]).min() >= 0.2 - 1e-4  # the rows' own tolerance (1e-5 x lever)


def test_terminal_stage_has_its_own_bounds_and_weight():
    """The planner's differences from the controller's NLP, one at a time, on the small case (dense oracle beside the host
    build): no terminal box -> the end state leaves the box; terminal weight 1 instead of 100 -> a different objective."""
    c = next(x for x in CASES if x["name"] == "plan_n10_k100_2obs")
    N = int(c["horizon"])
    cfg = planner_preset(N)
    obs = Obstacles.from_list([tuple(r) for r in c["rects"]])
    obst = [dict(center=(r[0], r[1]), width=r[2], height=r[3]) for r in c["rects"]]
    for tw, tb in ((100.0, 0.0), (1.0, 1e-2)):
        r = emu.obca_plan_batch(cfg, obs, c["x_init"][None], c["goal"], tw, tb, guess_z(c)[None])
        nlp = ob.ObcaNlp(N, cfg.dt, cfg.L1, cfg.L2, cfg.M, 3.05, 2.95, cfg.Qm(), np.array(cfg.R[:]).reshape(2, 2), list(cfg.x_lb),
                         list(cfg.x_ub), list(cfg.u_lb), list(cfg.u_ub), obst, terminal_weight=tw,
                         terminal_box=(c["goal"], tb) if tb > 0 else None)
        d = ob.solve(nlp, c["x_init"], np.tile(c["goal"], (N + 1, 1)), np.zeros((N, 2)), tol=cfg.tol, acc_tol=cfg.acceptable_tol,
                     acc_iter=cfg.acceptable_iter, max_iter=500, guess=(c["guess"], np.zeros((N, 2))))
        assert r["status"][0] == 0 and d["status"] == 0
        xs, us = split_z(r["z"][0], N)
        assert np.abs(xs - d["states"]).max() <= Z_TOL and np.abs(us - d["inputs"]).max() <= Z_TOL
        if tb == 0.0:  # a relaxation of the planner's problem: cheaper, and the end state leaves the box
            assert r["obj"][0] < c["obj"] - 1e-2 and np.abs(xs[-1] - c["goal"]).max() > TERM_BOX
        else:
            assert abs(r["obj"][0] - c["obj"]) > 1e-2  # the terminal term is weighted differently


def test_initial_trajectory_of_the_shim_follows_the_reference():
    """trajectory_optimization.py:227-274: spline through the waypoints sampled at `horizon` nodes, headings + pi/2,
    steering angle / speed / inputs zero, the last node repeated; two waypoints (the repository's initialize.json) give a
    straight line."""
    from car_trailer_mpc_b200.trajectory_optimization import TrajectoryOptimization, interpolate_waypoints
    t = object.__new__(TrajectoryOptimization)
    t._horizon = 20
    wp = {"Positions": [[38.5, 26.0], [15.5, 12.45]], "Headings": [-1.309, 0.0], "HitchAngles": [0.0, 0.0]}  # initialize.json
    g = t._hybrid_a_star_initial_trajectory(wp)
    assert g.shape == (21, 6) and np.array_equal(g[20], g[19]) and np.all(g[:, 4:] == 0)
    assert np.allclose(g[0, :3], [38.5, 26.0, -1.309 + np.pi / 2]) and np.allclose(g[19, :3], [15.5, 12.45, np.pi / 2])
    assert np.allclose(g[:20, 0], np.linspace(38.5, 15.5, 20))
    from scipy.interpolate import CubicSpline
    w3 = np.array([[0.0, 0.0], [1.0, 2.0], [3.0, 1.0], [4.0, 4.0]])
    assert np.allclose(interpolate_waypoints(w3, 9), CubicSpline(np.linspace(0, 1, 4), w3)(np.linspace(0, 1, 9)))
    s = t._generate_initial_trajectory_guess(np.arange(6.0), np.ones(6))
    assert np.allclose(s[0], np.arange(6.0)) and np.allclose(s[20], 1.0) and np.allclose(s[10], 0.5 * (np.arange(6.0) + 1.0))
