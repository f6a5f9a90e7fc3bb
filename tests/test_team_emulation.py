"""The warp-cooperative ("team") solve (csrc/ttmpc_team.cuh: L lanes per problem, iterate in shared memory, stage-parallel
evaluation + cooperative Riccati / forward / costate recursions), compiled for the host by tools/team_emu.cpp -- the 32
lanes of a warp run as fibers -- must agree with the oracle AND with the lane-per-problem core (tools/kernel_emu.cpp):
three independently structured implementations of the same algorithm.  Runs without a GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emu  # noqa: E402
from parity import assert_parity  # noqa: E402

from car_trailer_mpc_b200 import nmpc_preset, tracking_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402
from oracle import oracle  # noqa: E402


@pytest.mark.parametrize("lanes", [8, 16, 32])
def test_team_matches_oracle_narrow_and_wide(lanes):
    cfg = tracking_preset(40); cfg.max_iter = 200
    for sig in (pb.SIGMA_NARROW, pb.SIGMA_WIDE):
        sc = pb.make_scenarios(cfg, 96, sigma=sig)
        r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, nthreads=4)
        r1 = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=lanes)
        assert_parity(cfg, r1, r0, sc.x_init)
        assert (r0["iters"] == r1["iters"]).mean() > 0.98


@pytest.mark.parametrize("horizon,lanes", [(1, 8), (7, 8), (10, 8), (15, 16), (16, 16), (31, 32), (32, 32), (73, 32), (128, 32), (100, 16)])
def test_team_other_horizons(horizon, lanes):
    """Horizons around the pass boundaries (N + 1 = k * lanes and +-1), the shortest and the longest."""
    cfg = tracking_preset(horizon)
    sc = pb.make_scenarios(cfg, 13, seed=5 + horizon)
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, nthreads=4)
    r1 = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=lanes)
    assert_parity(cfg, r1, r0, sc.x_init)


def test_team_nmpc_preset_warm_start_and_general_weights():
    cfg = nmpc_preset(30)
    sc = pb.make_scenarios(cfg, 24, seed=5)
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    r1 = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=16)
    assert np.array_equal(r0["status"], r1["status"])          # tol 1e-3: both stop within ~1e-3 of the KKT point
    assert np.abs(r0["u0"] - r1["u0"]).max() < 1e-6
    Q = np.diag([1.0, 1.0, 2.0, 3.0, 1.0, 1.0]); Q[0, 1] = Q[1, 0] = 0.3; Q[2, 5] = Q[5, 2] = -0.2
    R = np.array([[5.0, 0.7], [0.7, 8.0]])
    cfg.set_weights(Q, R); cfg.tol = 1e-8; cfg.acceptable_tol = 1e-6; cfg.acceptable_iter = 15
    sc = pb.make_scenarios(cfg, 24, seed=9, families=False)
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    zw = pb.shift_warm_start(r0["z"], 30, reference_bug=True)
    r0w = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, z_warm=zw)
    for lanes in (8, 16):
        r1w = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, z_warm=zw, lanes=lanes)
        assert_parity(cfg, r1w, r0w, sc.x_init)
        assert_parity(cfg, r1w, r0, sc.x_init)   # warm and cold starts reach the same minimiser


def test_team_infeasible_x0_policy():
    cfg = tracking_preset(20)
    sc = pb.make_scenarios(cfg, 5, seed=2, families=False)
    sc.x_init[1, 4] = 0.9   # phi beyond pi/4: reference NLP infeasible
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    r1 = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=8)
    assert r0["status"][1] == 5 and r1["status"][1] == 5
    assert np.array_equal(r0["status"], r1["status"])


def test_team_generic_bounds_and_dense_weight_variants():
    cfg = tracking_preset(40)
    sc = pb.make_scenarios(cfg, 32, seed=13, sigma=pb.SIGMA_WIDE)
    a = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=16)
    for flags in (1, 2, 3):   # generic-bounds / dense-weights code paths on the default problem: same arithmetic
        b = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=16, flags=flags)
        assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["iters"], b["iters"]), flags
        assert np.abs(a["z"] - b["z"]).max() < 1e-9, flags
    # general masks: x,y bounded, theta free, v one-sided
    cfg.set_bounds([-200.0, -200.0, -np.inf, -1.0, -0.7, -np.inf], [200.0, 200.0, np.inf, 1.0, 0.7, 6.0], [-4.0, -1.0], [4.0, 1.0])
    sc = pb.make_scenarios(cfg, 32, seed=14, families=False)
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    r1 = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=16)
    assert_parity(cfg, r1, r0, sc.x_init)


def test_team_shared_trajectory_mode(traj):
    S, U = traj
    cfg = tracking_preset(40)
    rng = np.random.default_rng(5)
    k = np.concatenate([rng.integers(0, 460, size=17), [0, 360, 361, 399, 400, 401, 1000]]).astype(np.int32)
    xs, us = pb.windows_batch(S, U, k, 40)
    x = xs[:, 0, :] + rng.normal(0, 0.02, size=(len(k), 6))
    a = emu.team_solve_batch(cfg, x, xs, us, lanes=16)
    b = emu.team_solve_batch(cfg, x, k_index=k, traj_states=S, traj_inputs=U, lanes=16)
    for key in ("z", "u0", "obj", "iters", "status"):
        assert np.array_equal(a[key], b[key]), key


def test_team_far_off_batch_walks_the_lane_kernels_iterates():
    """Far-off initial states: backtracking, line-search failures, inertia corrections and infeasible x_0.  The team
    kernel must end every problem with the status of the lane-per-problem core and, where that converged, at its
    minimiser.  (Iterates differ in the last bits: statistics are summed in a different order.)"""
    cfg = tracking_preset(30); cfg.max_iter = 60
    sc = pb.make_scenarios(cfg, 64, seed=17, sigma=pb.SIGMA_WIDE * 3.0)
    a = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    b = emu.team_solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, lanes=16)
    assert (a["iters"] > 12).any() and (a["status"] != 0).any()
    same = a["status"] == b["status"]
    assert same.mean() >= 0.95, (a["status"], b["status"])
    conv = same & (a["status"] == 0)
    assert np.abs(a["u0"][conv] - b["u0"][conv]).max() < 1e-6
    assert (a["iters"][conv] == b["iters"][conv]).mean() > 0.9
