"""Obstacle-aware controller on the GPU, called through the C ABI
(ttmpc_obca_solve_batch[_shared]): against the dense oracle's golden solves, against the host build of the same solver
core on seeded batches, through the MPCTrackingControlObs shim, and -- at config 4's full size -- through
size-independent properties (dynamics defect, true geometric clearance >= d_min, plain tracking solve is a relaxation)."""
import os
import sys

import numpy as np
import pytest

import geometry
from obca_common import GOLD_OBCA_FULL, OBCA_FLAVOURS, Z_TOL, case_problem, force_obca_kernel, golden_cases, split_z
from parity import OBJ_REL_TOL, U0_ABS_TOL, VIOL_TOL

from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200 import tracking_preset
from car_trailer_mpc_b200.config import Obstacles, parking_lot_obstacles

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
CASES = golden_cases()
FULL = golden_cases(GOLD_OBCA_FULL)
LOT = [(o["center"][0], o["center"][1], o["width"], o["height"]) for o in parking_lot_obstacles()]


def solver(cfg):
    from car_trailer_mpc_b200 import BatchSolver
    return BatchSolver(cfg, 0)


def scenarios(cfg, B, seed, kmax=330, sigma=0.002):
    S, U = pb.load_reference_trajectory()
    rng = np.random.default_rng(seed)
    ks = rng.integers(0, kmax + 1, B).astype(np.int32)
    x0 = S[ks] + rng.normal(0, sigma, (B, 6))
    lb, ub = np.array(cfg.x_lb[:]), np.array(cfg.x_ub[:])
    x0[:, 2:] = np.clip(x0[:, 2:], lb[2:] + 1e-3, ub[2:] - 1e-3)
    return S, U, ks, x0


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_gpu_matches_dense_oracle_golden(c):
    cfg, obs = case_problem(c)
    r = solver(cfg).solve_obca(obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])  # host pointers
    assert r["status"][0] == 0
    xs, us = split_z(r["z"][0], cfg.horizon)
    assert np.abs(r["u0"][0] - c["inputs"][0]).max() <= U0_ABS_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    assert np.array_equal(xs[0], c["x_init"])


@pytest.mark.parametrize("flavour", OBCA_FLAVOURS)
@pytest.mark.parametrize("c", FULL, ids=[c["name"] for c in FULL])
def test_gpu_matches_oracle_at_the_reference_size(c, flavour, monkeypatch):
    """Config 4's own size (simulation.py:390: horizon 50 / 40, the 11 rectangles of obstacles.json, one case with a twelfth
    obstacle and active rows, one in which the oracle recovers from an exhausted line search): both GPU kernels, through
    the C ABI, against the oracle's block-tridiagonal LDL' solutions (tools/make_golden_obca_full.py)."""
    kernel = force_obca_kernel(monkeypatch, flavour)
    cfg, obs = case_problem(c)
    cfg.max_iter = 400
    sv = solver(cfg)
    r = sv.solve_obca(obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert sv.kernel_launches()[kernel] == 1
    assert r["status"][0] == 0
    xs, us = split_z(r["z"][0], cfg.horizon)
    assert np.abs(r["u0"][0] - c["inputs"][0]).max() <= U0_ABS_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    assert r["kkt"][0][1] <= VIOL_TOL


@pytest.mark.parametrize("flavour", OBCA_FLAVOURS)
def test_gpu_matches_host_build_of_the_core_on_a_seeded_batch(flavour, monkeypatch):
    """Both kernels of the obstacle-aware path: TTMPC_OBCA_WIDE_MAX=0 forces ttmpc_obca_kernel (one warp per problem),
    the default takes ttmpc_obca_wide_kernel (one CTA per problem) for a batch of this size."""
    import emu
    import torch
    kernel = force_obca_kernel(monkeypatch, flavour)
    cfg = tracking_preset(20)
    cfg.max_iter = 300
    obs = Obstacles.from_list(parking_lot_obstacles())
    S, U, ks, x0 = scenarios(cfg, 40, seed=7, kmax=400)  # the tail (k > 340) rides the d_min boundary: failures included
    dev = torch.device("cuda:0")
    sv = solver(cfg)
    g = sv.solve_obca_shared(obs, torch.from_numpy(x0).to(dev), torch.from_numpy(ks).to(dev),
                             torch.from_numpy(S).to(dev), torch.from_numpy(U).to(dev))
    g = {k: v.cpu().numpy() for k, v in g.items()}
    launches = sv.kernel_launches()
    assert launches[kernel] == 1
    e = emu.obca_solve_batch(cfg, obs, x0, k_index=ks, traj_states=S, traj_inputs=U)
    ok = (g["status"] == 0) & (e["status"] == 0)
    assert ok.sum() >= 30
    assert ((g["status"] <= 1) == (e["status"] <= 1)).mean() >= 0.95  # borderline instances may flip
    assert np.abs(g["u0"][ok] - e["u0"][ok]).max() <= U0_ABS_TOL
    assert (np.abs(g["obj"][ok] - e["obj"][ok]) / np.maximum(1e-12, np.abs(e["obj"][ok]))).max() <= OBJ_REL_TOL
    assert np.abs(g["z"][ok] - e["z"][ok]).max() <= Z_TOL


def test_shim_is_a_drop_in(capsys):
    """MPCTrackingControlObs(dynamics, params, Q, R, state_bound, input_bound, obstacle_list).solve(...) as
    simulation.py:416-429,520 uses it."""
    from car_trailer_mpc_b200 import MPCTrackingControlObs, TruckTrailerModel
    c = next(x for x in CASES if x["name"] == "n12_k200_blocked3")
    N = int(c["horizon"])
    params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05, "horizon": N}
    pi = np.pi
    sb = {"lb": [-np.inf, -np.inf, -pi, -pi / 3, -pi / 4, -10.0], "ub": [np.inf, np.inf, pi, pi / 3, pi / 4, 10.0]}
    ib = {"lb": [-5.0, -pi / 2], "ub": [5.0, pi / 2]}
    obst = [dict(center=(r[0], r[1]), width=r[2], height=r[3]) for r in c["rects"]]
    ctl = MPCTrackingControlObs(TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib, obstacle_list=obst)
    states, inputs = ctl.solve(c["x_init"], c["ref_states"].T.copy(), c["ref_inputs"].T.copy())
    assert states.shape == (6, N + 1) and inputs.shape == (2, N)
    assert np.abs(states.T - c["states"]).max() <= Z_TOL and np.abs(inputs.T - c["inputs"]).max() <= Z_TOL
    assert ctl.last_status == 0 and "Cannot find" not in capsys.readouterr().out
    # no obstacles -> the plain tracking NLP (mpc_control_obs.py:181-188)
    from car_trailer_mpc_b200 import MPCTrackingControl
    a = MPCTrackingControlObs(TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib, obstacle_list=[])
    b = MPCTrackingControl(TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib)
    sa, ia = a.solve(c["x_init"], c["ref_states"].T.copy(), c["ref_inputs"].T.copy())
    sb_, ib_ = b.solve(c["x_init"], c["ref_states"].T.copy(), c["ref_inputs"].T.copy())
    assert np.array_equal(sa, sb_) and np.array_equal(ia, ib_)


def test_full_size_properties_config4():
    """N = 50, all 11 rectangles of obstacles.json, 2048 seeded scenarios: every converged solution satisfies the
    dynamics, the bounds and the TRUE body-obstacle distance >= d_min; its cost is >= the plain tracking optimum."""
    import torch
    cfg = tracking_preset(50)
    cfg.max_iter = 300
    obs = Obstacles.from_list(parking_lot_obstacles())
    B = 2048
    S, U, ks, x0 = scenarios(cfg, B, seed=20251018, kmax=320)
    dev = torch.device("cuda:0")
    s = solver(cfg)
    args = (torch.from_numpy(x0).to(dev), torch.from_numpy(ks).to(dev), torch.from_numpy(S).to(dev), torch.from_numpy(U).to(dev))
    g = {k: v.cpu().numpy() for k, v in s.solve_obca_shared(obs, *args).items()}
    plain = {k: v.cpu().numpy() for k, v in s.solve_shared(*args).items()}
    ok = g["status"] <= 1
    # The remainder are windows through the tight passage around k = 117 (trailer 0.215 m from a corner at 4.9 m/s): the
    # iteration jams against the bounds with theta ~ 1 -- where Ipopt switches to its restoration phase, which neither
    # the oracle nor the kernel restates (both report status 3; tests/test_obca_cpu.py pins that they agree).
    assert ok.mean() >= 0.92, np.bincount(g["status"], minlength=6)
    assert ok[(ks < 60) | ((ks > 140) & (ks < 290))].mean() >= 0.97
    X, Uu = split_z(g["z"][ok], 50)
    assert np.abs(pb.dynamics_defect(cfg, X, Uu)).max() <= VIOL_TOL
    assert np.array_equal(X[:, 0], x0[ok])
    lb, ub = np.array(cfg.x_lb[:]), np.array(cfg.x_ub[:])
    assert (X[:, 1:] >= lb - 2e-7).all() and (X[:, 1:] <= ub + 2e-7).all()
    cl = geometry.clearance(X, LOT)
    assert cl.min() >= 0.2 - 1e-4, cl.min()
    both = ok & (plain["status"] <= 1)
    assert (g["obj"][both] >= plain["obj"][both] * (1 - 1e-6) - 1e-9).all()
    # far from every obstacle the two controllers agree
    far = both & (geometry.clearance(S[np.minimum(ks[:, None] + np.arange(51), 400)], LOT).min(1) > 1.0)
    assert far.sum() > 100
    assert np.abs(g["u0"][far] - plain["u0"][far]).max() <= U0_ABS_TOL


def test_closed_loop_with_the_obstacle_aware_controller():
    """simulation.py:484-531 with USE_OBS_MPC: the headless loop drives MPCTrackingControlObs for 2 s of the manoeuvre
    (40 control steps: the float-accumulated clock of simulation.py:484 stops short of t = 2.0).  Far from the parking-lot rectangles the collision rows are inactive, so the loop must
    reproduce the loop run with the plain controller; every solve succeeds and the path keeps its distance."""
    from car_trailer_mpc_b200 import MPCTrackingControl, MPCTrackingControlObs, TruckTrailerModel
    from car_trailer_mpc_b200 import closed_loop as cl
    S, U = pb.load_reference_trajectory()
    N = 20
    params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05, "horizon": N}
    pi = np.pi
    sb = {"lb": [-np.inf, -np.inf, -pi, -pi / 3, -pi / 4, -10.0], "ub": [np.inf, np.inf, pi, pi / 3, pi / 4, 10.0]}
    ib = {"lb": [-5.0, -pi / 2], "ub": [5.0, pi / 2]}
    args = (TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib)
    obs_ctl = MPCTrackingControlObs(*args, obstacle_list=parking_lot_obstacles())
    plain_ctl = MPCTrackingControl(*args)
    x0 = S[0] + np.array([0.3, -0.2, 0.02, 0.0, 0.0, 0.0])
    a = cl.simulate_single(obs_ctl, S, U, x0, 2.0, 0.05, N, params)
    b = cl.simulate_single(plain_ctl, S, U, x0, 2.0, 0.05, N, params)
    assert b.failures == 0 and len(a.controls) == len(b.controls) >= 40
    # The cold start of the reference (all OBCA duals at 100, mpc_control_obs.py:226-237) makes a few of these solves crawl
    # for 100-200 iterations under heavy regularisation before they converge, and an occasional one ends as a line-search
    # failure (where Ipopt would enter restoration); the loop then applies the last iterate, as the reference does.
    assert a.failures <= 2
    if a.failures == 0:
        assert np.abs(a.controls - b.controls).max() <= U0_ABS_TOL
        assert np.abs(a.states - b.states).max() <= 1e-5
    else:  # identical up to the first failed solve
        first = int(np.argmax(np.abs(a.controls - b.controls).max(1) > U0_ABS_TOL))
        assert first >= 5 and np.abs(a.controls[:first] - b.controls[:first]).max() <= U0_ABS_TOL
    assert geometry.clearance(a.states, LOT).min() >= 0.2


def test_closed_loop_with_the_geometric_start():
    """The same 2 s of the manoeuvre with ``MPCTrackingControlObs(..., geometric_start=True)`` (opt-in, not the reference's
    dual guesses): no solve crawls or fails, every solve takes a fraction of the iterations, and -- the collision rows being
    inactive out here -- the loop equals the plain controller's."""
    from car_trailer_mpc_b200 import MPCTrackingControl, MPCTrackingControlObs, TruckTrailerModel
    from car_trailer_mpc_b200 import closed_loop as cl
    S, U = pb.load_reference_trajectory()
    N = 20
    params = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05, "horizon": N}
    pi = np.pi
    sb = {"lb": [-np.inf, -np.inf, -pi, -pi / 3, -pi / 4, -10.0], "ub": [np.inf, np.inf, pi, pi / 3, pi / 4, 10.0]}
    ib = {"lb": [-5.0, -pi / 2], "ub": [5.0, pi / 2]}
    args = (TruckTrailerModel(params), params, np.eye(6), 10.0 * np.eye(2), sb, ib)
    geo_ctl = MPCTrackingControlObs(*args, obstacle_list=parking_lot_obstacles(), geometric_start=True)
    plain_ctl = MPCTrackingControl(*args)
    x0 = S[0] + np.array([0.3, -0.2, 0.02, 0.0, 0.0, 0.0])
    iters = []
    solve = geo_ctl.solve

    def counting_solve(*a, **k):
        out = solve(*a, **k)
        iters.append(geo_ctl.last_iterations)
        return out

    geo_ctl.solve = counting_solve
    a = cl.simulate_single(geo_ctl, S, U, x0, 2.0, 0.05, N, params)
    b = cl.simulate_single(plain_ctl, S, U, x0, 2.0, 0.05, N, params)
    assert a.failures == 0 and b.failures == 0 and len(a.controls) == len(b.controls) >= 40
    assert max(iters) <= 40  # the reference start needs 20 .. 200 on these solves
    assert np.abs(a.controls - b.controls).max() <= U0_ABS_TOL
    assert np.abs(a.states - b.states).max() <= 1e-5
    assert geometry.clearance(a.states, LOT).min() >= 0.2


def test_obca_edge_cases():
    """Empty batch, the shortest horizon, the maximum obstacle count (16 -> all 32 lanes of the warp busy), a batch that
    is not a multiple of the CTA's 8 problem slots, and a rejected obstacle set."""
    import emu
    from car_trailer_mpc_b200 import _lib
    S, U = pb.load_reference_trajectory()
    lot16 = parking_lot_obstacles() + [dict(center=(100.0 + 10 * i, -40.0), width=4.0, height=4.0) for i in range(5)]
    for N, B, obst in ((1, 3, parking_lot_obstacles()), (7, 13, lot16)):
        cfg = tracking_preset(N)
        cfg.max_iter = 300
        s = solver(cfg)
        obs = Obstacles.from_list(obst)
        rng = np.random.default_rng(N)
        ks = rng.integers(0, 300, B).astype(np.int32)
        xs, us = pb.windows_batch(S, U, ks, N)
        x0 = xs[:, 0] + rng.normal(0, 0.002, (B, 6))
        g = s.solve_obca(obs, x0, xs, us)
        e = emu.obca_solve_batch(cfg, obs, x0, xs, us)
        assert np.array_equal(g["status"], e["status"]) and (g["status"] == 0).all()
        assert np.abs(g["z"] - e["z"]).max() <= Z_TOL
        empty = s.solve_obca(obs, np.zeros((0, 6)), np.zeros((0, N + 1, 6)), np.zeros((0, N, 2)))
        assert empty["u0"].shape == (0, 2)
    bad = Obstacles.from_list([(0.0, 0.0, 0.0, 1.0)])  # zero width
    with pytest.raises(_lib.TTMPCError):
        s.solve_obca(bad, x0, xs, us)
    with pytest.raises(ValueError):
        Obstacles.from_list([(0.0, 0.0, 1.0, 1.0)] * 17)


@pytest.mark.parametrize("name", ["n6_k300_2obs", "n12_k60_blocked"])
def test_gpu_reaches_the_slsqp_minimiser_of_the_obstacle_aware_nlp(name):
    """The GPU's output against an algorithm that shares nothing with Ipopt's rules (tools/make_golden_slsqp_obca.py)."""
    g = np.load(os.path.join(os.path.dirname(GOLD_OBCA_FULL), "obca_slsqp.npz"))
    c = next(x for x in CASES if x["name"] == name)
    cfg, obs = case_problem(c)
    r = solver(cfg).solve_obca(obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert r["status"][0] == 0 and int(g[name + "/success"]) == 1
    xs, us = split_z(r["z"][0], cfg.horizon)
    assert np.abs(r["u0"][0] - g[name + "/inputs"][0]).max() <= U0_ABS_TOL
    assert np.abs(xs - g[name + "/states"]).max() <= 1e-5 and np.abs(us - g[name + "/inputs"]).max() <= 1e-5
    assert abs(r["obj"][0] - float(g[name + "/obj"])) <= 5e-6 * abs(r["obj"][0])


@pytest.mark.parametrize("flavour", OBCA_FLAVOURS)
def test_geometric_start_of_the_duals(flavour, monkeypatch):
    """TTMPC_OBCA_GEOMETRIC_START (opt-in, not the reference's starting point): on the golden cases at the reference's
    size every kernel converges in fewer iterations, to the golden solution or -- the NLP is not convex -- to another
    KKT point that is at least as good (lower objective, dynamics defect and true clearance checked); on a seeded batch
    the iteration count drops by more than half with no loss of converged instances."""
    import torch
    kernel = force_obca_kernel(monkeypatch, flavour)
    n_conv = 0
    for c in FULL + CASES:
        cfg, _ = case_problem(c)
        cfg.max_iter = 400
        rects = [tuple(r) for r in c["rects"]]
        sv = solver(cfg)
        ref = sv.solve_obca(Obstacles.from_list(rects), c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
        geo = sv.solve_obca(Obstacles.from_list(rects, geometric_start=True), c["x_init"][None], c["ref_states"][None],
                            c["ref_inputs"][None])
        assert sv.kernel_launches()[kernel] == 2
        if geo["status"][0] != 0:
            # not uniformly better: the case with a twelfth obstacle across the path (active rows everywhere) heads for a
            # cheaper region from this start and jams in the line search -- the one known exception among the goldens
            assert c["name"] == "n50_k60_12obs_blocked" and geo["status"][0] == 3
            continue
        n_conv += 1
        assert geo["iters"][0] < ref["iters"][0]
        xs, us = split_z(geo["z"][0], cfg.horizon)
        same = np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
        assert same or geo["obj"][0] <= c["obj"] * (1 + 1e-9)
        assert np.abs(pb.dynamics_defect(cfg, xs[None], us[None])).max() <= 1e-6
        assert geometry.clearance(xs, rects).min() >= 0.2 - 1e-4
    assert n_conv >= len(FULL + CASES) - 1
    cfg = tracking_preset(30)
    cfg.max_iter = 300
    S, U, ks, x0 = scenarios(cfg, 64, seed=11)
    dev = torch.device("cuda:0")
    sv = solver(cfg)
    args = (torch.from_numpy(x0).to(dev), torch.from_numpy(ks).to(dev), torch.from_numpy(S).to(dev), torch.from_numpy(U).to(dev))
    ref = sv.solve_obca_shared(Obstacles.from_list(parking_lot_obstacles()), *args)
    geo = sv.solve_obca_shared(Obstacles.from_list(parking_lot_obstacles(), geometric_start=True), *args)
    rs, gs = ref["status"].cpu().numpy(), geo["status"].cpu().numpy()
    assert (gs <= 1).sum() >= (rs <= 1).sum()
    assert geo["iters"].float().mean().item() < 0.5 * ref["iters"].float().mean().item()
    both = (rs == 0) & (gs == 0)
    dz = (geo["z"] - ref["z"]).abs().amax(dim=1).cpu().numpy()
    better = (geo["obj"] <= ref["obj"] * (1 + 1e-9)).cpu().numpy()
    assert ((dz <= Z_TOL) | better)[both].mean() >= 0.9  # same KKT point, or a better one, on at least nine out of ten


@pytest.mark.parametrize("size", ["2", "4", "16"])
def test_cluster_sizes(size, monkeypatch):
    """The cluster-per-problem kernel at the other cluster sizes (the flavour tests force 8; 16 is the non-portable maximum
    and falls back to 8 on a device that does not take it): a golden case at the reference's size, and a small batch that
    makes the persistent clusters pull several problems each from the queue."""
    import torch
    c = next(x for x in FULL if x["name"] == "n50_k200_11obs")
    kernel = force_obca_kernel(monkeypatch, "cluster_per_problem", cluster=size)
    cfg, obs = case_problem(c)
    cfg.max_iter = 400
    sv = solver(cfg)
    r = sv.solve_obca(obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert sv.kernel_launches()[kernel] == 1
    assert r["status"][0] == 0
    xs, us = split_z(r["z"][0], cfg.horizon)
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    # 24 problems, horizon 20: more problems than clusters fit at sizes 8 / 16 -- same results as one CTA per problem
    cfg = tracking_preset(20)
    cfg.max_iter = 300
    S, U, ks, x0 = scenarios(cfg, 24, seed=5)
    dev = torch.device("cuda:0")
    args = (torch.from_numpy(x0).to(dev), torch.from_numpy(ks).to(dev), torch.from_numpy(S).to(dev), torch.from_numpy(U).to(dev))
    sv = solver(cfg)
    a = sv.solve_obca_shared(Obstacles.from_list(parking_lot_obstacles()), *args)
    assert sv.kernel_launches()[kernel] == 1
    monkeypatch.setenv("TTMPC_OBCA_CLUSTER", "0")
    b = sv.solve_obca_shared(Obstacles.from_list(parking_lot_obstacles()), *args)
    assert sv.kernel_launches()["ttmpc_obca_wide_kernel"] == 1
    assert torch.equal(a["status"], b["status"]) and torch.equal(a["iters"], b["iters"])  # same arithmetic, same order
    assert torch.equal(a["z"], b["z"])
