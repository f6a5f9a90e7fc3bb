"""Offline planner on the GPU: ttmpc_plan_batch through the C ABI against the oracle's golden solutions of
the planner's NLP (tools/make_golden_planner.py; the last case is the reference's own size: horizon 200, 11 rectangles,
37 k variables, one CTA), both kernels, and the TrajectoryOptimization shim as trajectory_animation.py:81-109 uses it."""
import os

import numpy as np
import pytest

import geometry
from obca_common import OBCA_FLAVOURS, Z_TOL, force_obca_kernel, golden_cases, split_z
from parity import OBJ_REL_TOL, U0_ABS_TOL

from car_trailer_mpc_b200 import planner_preset
from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200.config import Obstacles, parking_lot_obstacles

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "planner_cases.npz")
CASES = golden_cases(GOLD)


def guess_z(c):
    N = int(c["horizon"])
    z = np.zeros(8 * N + 6)
    for k in range(N + 1):
        z[8 * k:8 * k + 6] = c["guess"][k]
    return z


@pytest.mark.parametrize("flavour", OBCA_FLAVOURS)
@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_gpu_matches_oracle_on_the_planner_nlp(c, flavour, monkeypatch):
    from car_trailer_mpc_b200 import BatchSolver
    kernel = force_obca_kernel(monkeypatch, flavour)
    N = int(c["horizon"])
    cfg = planner_preset(N)
    cfg.max_iter = 1000
    sv = BatchSolver(cfg, 0)
    obs = Obstacles.from_list([tuple(r) for r in c["rects"]])
    r = sv.plan(obs, c["x_init"][None], c["goal"], 100.0, 1e-2, guess_z(c)[None])  # host pointers
    assert sv.kernel_launches()[kernel] == 1
    assert r["status"][0] == 0
    xs, us = split_z(r["z"][0], N)
    assert np.abs(r["u0"][0] - c["inputs"][0]).max() <= U0_ABS_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    assert np.abs(pb.dynamics_defect(cfg, xs[None], us[None])).max() <= 1e-8
    assert (np.abs(xs[-1] - c["goal"]) <= 1e-2 + 1.01e-8 * np.maximum(1.0, np.abs(c["goal"]) + 1e-2)).all()
    assert geometry.clearance(xs, [tuple(q) for q in c["rects"]]).min() >= 0.2 - 1e-4


def test_device_pointer_batch_equals_host_pointer_calls():
    """B start states towards one goal on device arrays (torch) == the same problems one by one on host arrays."""
    import torch
    from car_trailer_mpc_b200 import BatchSolver
    c = next(x for x in CASES if x["name"] == "plan_n40_k150_11obs")
    N = int(c["horizon"])
    cfg = planner_preset(N)
    cfg.max_iter = 1000
    sv = BatchSolver(cfg, 0)
    obs = Obstacles.from_list([tuple(r) for r in c["rects"]])
    rng = np.random.default_rng(2)
    x0 = c["x_init"][None] + np.concatenate([np.zeros((1, 6)), rng.normal(0, 1e-3, (3, 6))])
    zg = np.tile(guess_z(c), (4, 1))
    dev = torch.device("cuda:0")
    d = sv.plan(obs, torch.from_numpy(x0).to(dev), c["goal"], 100.0, 1e-2, torch.from_numpy(zg).to(dev))
    dz, dst = d["z"].cpu().numpy(), d["status"].cpu().numpy()
    assert dst[0] == 0
    for i in range(4):  # (a perturbed start next to an obstacle may end as a line-search failure: same outcome either way)
        h = sv.plan(obs, x0[i:i + 1], c["goal"], 100.0, 1e-2, zg[i:i + 1])
        assert h["status"][0] == dst[i] and np.array_equal(h["z"][0], dz[i])
    xs, us = split_z(dz[0], N)
    assert np.abs(xs - c["states"]).max() <= Z_TOL


def test_shim_is_a_drop_in_for_trajectory_optimization():
    """TrajectoryOptimization(model, params, Q, R, state_bound, input_bound, obstacle_list).plan(initial_state, goal_state)
    as trajectory_animation.py:41-110 drives it: horizon 200, dt 0.1, the 11 rectangles; waypoints = the stored poses (the
    reference reads a Hybrid-A* path from initialize.json)."""
    from car_trailer_mpc_b200 import TrajectoryOptimization, TruckTrailerModel
    S = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt")).T
    params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.1, "horizon": 200}
    pi, inf = np.pi, np.inf
    sb = {"lb": [-inf, -inf, -inf, -pi / 3, -pi / 4, -5.0], "ub": [inf, inf, inf, pi / 3, pi / 4, 10.0]}
    ib = {"lb": [-5.0, -pi / 2], "ub": [5.0, pi / 2]}
    planner = TrajectoryOptimization(TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib, parking_lot_obstacles())
    init, goal = S[0].copy(), S[200].copy()
    goal[4:] = 0.0
    wp = {"Positions": S[:, :2].tolist(), "Headings": (S[:, 2] - pi / 2).tolist(), "HitchAngles": S[:, 3].tolist()}
    states, inputs = planner.plan(init, goal, waypoints=wp)
    assert states.shape == (6, 201) and inputs.shape == (2, 200) and planner.converged()
    cfg = planner_preset(200)
    assert np.abs(pb.dynamics_defect(cfg, states.T[None], inputs.T[None])).max() <= 1e-8
    assert np.array_equal(states[:, 0], init) and np.abs(states[:, -1] - goal).max() <= 1e-2 + 1e-6
    rects = [(o["center"][0], o["center"][1], o["width"], o["height"]) for o in parking_lot_obstacles()]
    assert geometry.clearance(states.T, rects).min() >= 0.2 - 1e-4
    # the straight-line guess of _generate_initial_trajectory_guess (:208-225) cuts through the parked rows: whatever
    # the outcome, the call returns arrays of the right shape and an honest status
    s2, i2 = planner.plan(init, goal)
    assert s2.shape == (6, 201) and i2.shape == (2, 200) and planner.last_status in range(6)
