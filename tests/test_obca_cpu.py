"""Obstacle-aware controller, CPU side:
  * the dense numpy oracle (oracle/obca_oracle.py): derivatives by finite differences, golden solves reproduce;
  * the kernel's solver core compiled for the host (tools/obca_emu.cpp: pair condensation + Riccati) against the dense
    oracle -- two independent linear-algebra paths for the same algorithm;
  * the dual OBCA rows against an independent geometric distance computation (tests/geometry.py).
The reference (CasADi -> Ipopt) cannot run in this image: parity is unpinned, see the oracle's header."""
import json
import os
import sys

import numpy as np
import pytest

import geometry
from obca_common import GOLD_OBCA_FULL, Z_TOL, case_problem, golden_cases, split_z
from parity import OBJ_REL_TOL, U0_ABS_TOL

from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200 import tracking_preset
from car_trailer_mpc_b200.config import Obstacles, parking_lot_obstacles
from oracle import obca_oracle as ob

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emu  # noqa: E402

CASES = golden_cases()


def make_nlp(cfg, rects, W1=3.05, W2=2.95):
    obst = [dict(center=(r[0], r[1]), width=r[2], height=r[3]) for r in rects]
    return ob.ObcaNlp(cfg.horizon, cfg.dt, cfg.L1, cfg.L2, cfg.M, W1, W2, cfg.Qm(), np.array(cfg.R[:]).reshape(2, 2),
                      list(cfg.x_lb), list(cfg.x_ub), list(cfg.u_lb), list(cfg.u_ub), obst)


def test_parking_lot_matches_reference_file():
    lot = parking_lot_obstacles()
    assert len(lot) == 11
    assert lot[0] == dict(center=(-15.0, 10.0), width=30.0, height=20.0)
    assert lot[4] == dict(center=(21.5, 10.0), width=5.0, height=20.0)
    ref = "/root/reference/python-files/obstacles.json"
    if os.path.exists(ref):  # get_obstacles.py:5-33 restated on the reference's own file (absent on the GPU box)
        data = json.load(open(ref))
        for o, d in zip(lot, data):
            xs = [d[c]["X"] for c in ("FL", "FR", "BL", "BR")]
            ys = [d[c]["Y"] for c in ("FL", "FR", "BL", "BR")]
            assert o["center"] == (round(sum(xs) / 4, 4), round(sum(ys) / 4, 4))
            assert o["width"] == round(abs(d["FR"]["X"] - d["FL"]["X"]), 4)
            assert o["height"] == round(abs(d["BL"]["Y"] - d["FL"]["Y"]), 4)


def test_oracle_derivatives_finite_differences(traj):
    S, U = traj
    cfg = tracking_preset(3)
    rs, ru = pb.window(S, U, 300, 3)
    nlp = make_nlp(cfg, [(21.5, 10.0, 5.0, 20.0), (9.5, 10.0, 5.0, 20.0)])
    rng = np.random.default_rng(0)
    x0 = rs[0] + rng.normal(0, 0.02, 6)
    w = nlp.initial_point(x0, rs, ru) + 0.01 * rng.normal(0, 1, nlp.n)
    y = rng.normal(0, 1, nlp.m)
    h = 1e-6
    J = nlp.jacobian(w, x0)
    H = nlp.hessian(w, y, x0)
    Jfd = np.zeros_like(J)
    Hfd = np.zeros_like(H)

    def gl(v):
        return nlp.grad(v, rs, ru) + nlp.jacobian(v, x0).T @ y

    for i in range(nlp.n):
        e = np.zeros(nlp.n)
        e[i] = h
        Jfd[:, i] = (nlp.constraints(w + e, x0) - nlp.constraints(w - e, x0)) / (2 * h)
        Hfd[:, i] = (gl(w + e) - gl(w - e)) / (2 * h)
    assert np.abs(J - Jfd).max() <= 1e-6 * max(1.0, np.abs(J).max())
    assert np.abs(H - Hfd).max() <= 1e-6 * max(1.0, np.abs(H).max())
    assert np.abs(H - H.T).max() == 0.0


def test_dense_oracle_reproduces_golden_case():
    c = next(x for x in CASES if x["name"] == "n6_k300_2obs")
    cfg, _ = case_problem(c)
    r = ob.solve(make_nlp(cfg, c["rects"]), c["x_init"], c["ref_states"], c["ref_inputs"], tol=cfg.tol,
                 acc_tol=cfg.acceptable_tol, acc_iter=cfg.acceptable_iter, max_iter=cfg.max_iter)
    assert r["status"] == 0 and r["iters"] == int(c["iters"])
    assert np.abs(r["states"] - c["states"]).max() <= 1e-12 and np.abs(r["inputs"] - c["inputs"]).max() <= 1e-12


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_kernel_core_matches_dense_oracle(c):
    """Condensed (per-pair Cholesky + Riccati) vs dense LDL' linear algebra, same interior-point rules."""
    cfg, obs = case_problem(c)
    N = cfg.horizon
    r = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert r["status"][0] == 0
    xs, us = split_z(r["z"][0], N)
    assert np.abs(us[0] - c["inputs"][0]).max() <= U0_ABS_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    assert np.array_equal(xs[0], c["x_init"])
    # iteration counts: the OBCA duals have no cost, so the path through their flat directions depends on rounding
    assert abs(int(r["iters"][0]) - int(c["iters"])) <= max(10, int(c["iters"]) // 2)


@pytest.mark.parametrize("name", ["n12_k60_blocked", "n12_k200_blocked3"])
def test_blocking_obstacle_is_active_and_respected(name):
    """The OBCA rows certify distance >= d_min: measured with an independent polygon-distance routine the solution keeps
    exactly d_min to the blocking obstacle, while the reference window itself comes closer than d_min."""
    c = next(x for x in CASES if x["name"] == name)
    cfg, obs = case_problem(c)
    rects = [tuple(r) for r in c["rects"]]
    assert int(c["active_rows"]) >= 1
    assert geometry.clearance(c["ref_states"], rects).min() < 0.2 - 0.05
    cl = geometry.clearance(c["states"], rects)
    assert cl.min() >= 0.2 - 1e-4 and cl.min() <= 0.2 + 1e-4
    r = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    xs, _ = split_z(r["z"][0], cfg.horizon)
    cl = geometry.clearance(xs, rects)
    assert cl.min() >= 0.2 - 1e-4 and cl.min() <= 0.2 + 1e-4
    # the plain tracking solve of the same window is a relaxation: strictly cheaper
    plain = emu.solve_batch(cfg, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert plain["status"][0] == 0 and plain["obj"][0] < 0.5 * r["obj"][0]


def test_far_obstacles_do_not_change_the_tracking_solution(traj):
    """Inactive collision rows: states / inputs equal the plain controller's solution."""
    S, U = traj
    cfg = tracking_preset(20)
    rs, ru = pb.window(S, U, 30, 20)
    x0 = rs[0] + np.random.default_rng(3).normal(0, 0.02, 6)
    r = emu.obca_solve_batch(cfg, Obstacles.from_list(parking_lot_obstacles()), x0[None], rs[None], ru[None])
    plain = emu.solve_batch(cfg, x0[None], rs[None], ru[None])
    assert r["status"][0] == 0 and plain["status"][0] == 0
    assert np.abs(r["z"] - plain["z"]).max() <= 1e-6
    assert abs(r["obj"][0] - plain["obj"][0]) <= OBJ_REL_TOL * plain["obj"][0]


def test_start_inside_safety_margin_is_reported_infeasible(traj):
    """Stages 0 and 1 of the horizon are fixed by x_init; a start closer than d_min to an obstacle has no feasible
    point (Ipopt would end in restoration failure and the reference prints "Cannot find a solution!")."""
    S, U = traj
    cfg = tracking_preset(6)
    cfg.max_iter = 200
    rs, ru = pb.window(S, U, 240, 6)
    x, y, th = rs[0][:3]
    pc = np.array([x + np.cos(th) * 7.05 / 2, y + np.sin(th) * 7.05 / 2])
    n = np.array([-np.sin(th), np.cos(th)])
    c = pc + n * (3.05 / 2 + 0.05 + 0.5 * (abs(np.cos(th)) + abs(np.sin(th))))
    obs = Obstacles.from_list([(c[0], c[1], 1.0, 1.0)])
    r = emu.obca_solve_batch(cfg, obs, rs[0][None], rs[None], ru[None])
    assert r["status"][0] == 5 and r["iters"][0] == 30  # recognised from the geometry of stage 0, given up after 30 iterations
    d = ob.solve(make_nlp(cfg, [(c[0], c[1], 1.0, 1.0)]), rs[0], rs, ru, max_iter=200)
    assert d["status"] == 5 and d["iters"] == 30
    # the same obstacle 0.25 m away (> d_min): an ordinary solve
    c2 = pc + n * (3.05 / 2 + 0.25 + 0.5 * (abs(np.cos(th)) + abs(np.sin(th))))
    r = emu.obca_solve_batch(cfg, Obstacles.from_list([(c2[0], c2[1], 1.0, 1.0)]), rs[0][None], rs[None], ru[None])
    assert r["status"][0] == 0


def test_restored_duals_are_a_certificate_of_the_true_distance(traj):
    """Pair.restore (the recovery's closed-form OBCA duals for a fixed pose): its separation equals the polygon distance of
    an independent routine (tests/geometry.py), the duals are non-negative, rows c2 vanish, c1 = d_min - kappa * distance
    and c3 = kappa - 1 with kappa in (d_min / distance, 1) -- i.e. they prove distance >= d_min exactly when it holds."""
    S, _ = traj
    rng = np.random.default_rng(3)
    lot = parking_lot_obstacles()
    nlp = make_nlp(tracking_preset(4), [(o["center"][0], o["center"][1], o["width"], o["height"]) for o in lot])
    rects = [(o["center"][0], o["center"][1], o["width"], o["height"]) for o in lot]
    for _ in range(40):
        x = S[rng.integers(0, 401)] + rng.normal(0, 0.3, 6)
        veh, trl = geometry.body_corners(x)
        for j, pr in enumerate(nlp.pairs):
            box = geometry.box_corners(rects[j // 2])
            dist = float(geometry.poly_distance((veh, trl)[j % 2], box))
            v, sep = pr.restore(x[:4])
            if dist <= 0:  # overlapping: no separating direction
                assert sep <= 1e-12
                continue
            assert abs(sep - dist) <= 1e-9
            assert np.all(v >= 0)
            rows = pr.rows(x[:4], v)
            kappa = np.hypot(v[4] - v[6], v[5] - v[7])
            assert np.abs(rows[1:3]).max() <= 1e-12 and abs(rows[3] - (kappa - 1.0)) <= 1e-12
            assert abs(rows[0] - (ob.D_MIN - kappa * dist)) <= 1e-9
            if dist > ob.D_MIN:
                assert rows[0] < 0 and rows[3] < 0


FULL = golden_cases(GOLD_OBCA_FULL)


@pytest.mark.parametrize("c", FULL, ids=[c["name"] for c in FULL])
def test_kernel_core_matches_oracle_at_the_reference_size(c):
    """simulation.py:390's own size (horizon 50 / 40, the 11 rectangles of obstacles.json, 18 664 x 18 664 KKT matrix):
    per-pair condensation + Riccati (kernel core, host build) against the oracle's generic block-tridiagonal LDL'
    (tools/make_golden_obca_full.py).  One case has active collision rows; in one the oracle takes the recovery from an exhausted
    line search on its way to the same minimiser."""
    cfg, obs = case_problem(c)
    cfg.max_iter = 400
    N = cfg.horizon
    r = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert r["status"][0] == 0 and int(c["status"]) == 0
    xs, us = split_z(r["z"][0], N)
    assert np.abs(us[0] - c["inputs"][0]).max() <= U0_ABS_TOL
    assert abs(r["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
    assert np.abs(xs - c["states"]).max() <= Z_TOL and np.abs(us - c["inputs"]).max() <= Z_TOL
    # no assertion on iteration counts: with 1 122 cost-free dual blocks the path through their flat directions depends
    # on the last bits (30 ... 120 iterations between builds of the same source); the minimiser does not
    if "recovery" in c["name"]:
        assert int(c["restarts"]) >= 1  # the oracle's line search ran out of steps on the way and took the recovery
    if "blocked" in c["name"]:
        assert int(c["active_rows"]) >= 1
        rects = [tuple(q) for q in c["rects"]]
        assert abs(geometry.clearance(xs, rects).min() - 0.2) <= 1e-4  # rides d_min exactly


def test_banded_oracle_walks_the_dense_oracles_iterates():
    """linear_solver="banded" (what makes the reference's size reachable) is the same algorithm as the dense LDL'."""
    c = next(x for x in CASES if x["name"] == "n6_k300_2obs")
    cfg, _ = case_problem(c)
    r = ob.solve(make_nlp(cfg, c["rects"]), c["x_init"], c["ref_states"], c["ref_inputs"], tol=cfg.tol,
                 acc_tol=cfg.acceptable_tol, acc_iter=cfg.acceptable_iter, max_iter=cfg.max_iter, linear_solver="banded")
    assert r["status"] == 0 and r["iters"] == int(c["iters"])
    assert np.abs(r["states"] - c["states"]).max() <= 1e-12 and np.abs(r["inputs"] - c["inputs"]).max() <= 1e-12


def test_shared_trajectory_mode_equals_explicit_windows(traj):
    S, U = traj
    cfg = tracking_preset(10)
    obs = Obstacles.from_list(parking_lot_obstacles()[2:6])
    ks = np.array([5, 150, 395], np.int32)  # the last window runs past the end of the trajectory
    xs, us = pb.windows_batch(S, U, ks, 10)
    x0 = xs[:, 0] + np.random.default_rng(5).normal(0, 0.002, (3, 6))
    a = emu.obca_solve_batch(cfg, obs, x0, xs, us)
    b = emu.obca_solve_batch(cfg, obs, x0, k_index=ks, traj_states=S, traj_inputs=U)
    assert np.array_equal(a["status"], b["status"]) and np.array_equal(a["z"], b["z"])


def test_jammed_instance_fails_the_same_way_in_both_implementations(traj):
    """A window through the tight passage at k ~ 117 with only the nearest obstacle: the iteration jams (Ipopt would
    enter restoration, which is not restated).  The dense oracle and the kernel core must agree on that outcome."""
    S, U = traj
    lot = parking_lot_obstacles()
    N, k0 = 12, 110
    obst = sorted(lot, key=lambda o: abs(o["center"][0] - S[k0 + N // 2, 0]))[:1]
    cfg = tracking_preset(N)
    cfg.max_iter = 300
    rs, ru = pb.window(S, U, k0, N)
    rng = np.random.default_rng(0)
    for _ in range(3):
        rng.normal(0, 0.002, 6)
    x0 = rs[0] + rng.normal(0, 0.002, 6)
    e = emu.obca_solve_batch(cfg, Obstacles.from_list(obst), x0[None], rs[None], ru[None])
    rects = [(o["center"][0], o["center"][1], o["width"], o["height"]) for o in obst]
    d = ob.solve(make_nlp(cfg, rects), x0, rs, ru, max_iter=300)
    assert e["status"][0] == 3 and d["status"] == 3


def test_obca_core_under_address_sanitizer(tmp_path, traj):
    """compute-sanitizer is closed on the GPU pool: the scratch layout / pair indexing of the OBCA core run on the host
    under ASan + UBSan instead (5 problems > 3 emulated slots, so slots are reused)."""
    import ctypes
    import subprocess
    S, U = traj
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "obca_asan")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-o", exe,
                           os.path.join(root, "tools", "obca_asan_main.cpp"), os.path.join(root, "tools", "obca_emu.cpp"), "-lm"])
    N, B = 5, 5
    cfg = tracking_preset(N)
    cfg.max_iter = 60
    obs = Obstacles.from_list(parking_lot_obstacles()[:16])
    ks = np.array([10, 150, 250, 300, 398], np.int32)
    xs, us = pb.windows_batch(S, U, ks, N)
    x0 = xs[:, 0] + np.random.default_rng(9).normal(0, 0.002, (B, 6))
    open(tmp_path / "cfg.bin", "wb").write(bytes(ctypes.string_at(ctypes.byref(cfg), ctypes.sizeof(cfg))))
    open(tmp_path / "obs.bin", "wb").write(bytes(ctypes.string_at(ctypes.byref(obs), ctypes.sizeof(obs))))
    for name, arr in (("x", x0), ("xs", xs), ("us", us)):
        np.ascontiguousarray(arr, dtype=np.float64).tofile(tmp_path / f"{name}.bin")
    out = subprocess.run([exe, str(tmp_path), str(N), str(B)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "checksum 0 status" in out.stdout and "nan" not in out.stdout.lower()  # fused and wide runs agree exactly


@pytest.mark.parametrize("warps", [1, 3, 8])
def test_cta_per_problem_decomposition_is_bit_identical(warps):
    """ttmpc_obca_wide_kernel deals the stages of every sweep to the warps of a CTA and runs the recursions on warp 0;
    the host build of that decomposition must reproduce the single-warp sweeps exactly (same operations, other order
    of execution only; sums of statistics are taken per warp first)."""
    for c in CASES:
        cfg, obs = case_problem(c)
        cfg.max_iter = 300
        a = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
        b = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None], wide_warps=warps)
        assert a["status"][0] == b["status"][0] and a["iters"][0] == b["iters"][0]
        assert np.array_equal(a["z"], b["z"]) and np.array_equal(a["kkt"], b["kkt"])
        assert abs(a["obj"][0] - b["obj"][0]) <= 1e-14 * abs(a["obj"][0])  # the reported sum is taken in another order


# objective: SLSQP stops at the resolution of its finite-difference gradients on the case with active rows
SLSQP_OBJ_REL_TOL = 5e-6


@pytest.mark.parametrize("name", ["n6_k300_2obs", "n12_k60_blocked"])
def test_oracle_and_kernel_core_reach_the_slsqp_minimiser(name):
    """Independent algorithm for the obstacle-aware NLP: SciPy SLSQP on the literal inequality form of
    mpc_control_obs.py:65-176 (tools/make_golden_slsqp_obca.py: no slacks, no barrier, no code shared with the oracle).
    The dense oracle's golden solution and the kernel core (host build) must be that minimiser -- first control within the
    north-star tolerance, states and inputs within 1e-5; one of the two cases has active collision rows."""
    g = np.load(os.path.join(os.path.dirname(GOLD_OBCA_FULL), "obca_slsqp.npz"))
    assert int(g[name + "/success"]) == 1
    c = next(x for x in CASES if x["name"] == name)
    cfg, obs = case_problem(c)
    r = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    xs, us = split_z(r["z"][0], cfg.horizon)
    for X, U, J in ((c["states"], c["inputs"], float(c["obj"])), (xs, us, float(r["obj"][0]))):
        assert np.abs(U[0] - g[name + "/inputs"][0]).max() <= U0_ABS_TOL
        assert np.abs(X - g[name + "/states"]).max() <= 1e-5 and np.abs(U - g[name + "/inputs"]).max() <= 1e-5
        assert abs(J - float(g[name + "/obj"])) <= SLSQP_OBJ_REL_TOL * abs(J)


def test_geometric_start_of_the_duals_same_solution_fewer_iterations():
    """TTMPC_OBCA_GEOMETRIC_START (opt-in; NOT the reference's starting point): the OBCA duals start at the multipliers
    of the distance problems for the pose of the starting trajectory instead of mu = 100, lam = (100..115).  Oracle and
    host build of the kernel core implement the same rule, reach the golden solution of the reference start on a case
    away from the passage, and need a fraction of the iterations."""
    c = next(x for x in CASES if x["name"] == "n12_k200_blocked3")
    cfg, obs = case_problem(c)
    obs_geo = Obstacles.from_list([tuple(r) for r in c["rects"]], geometric_start=True)
    nlp = make_nlp(cfg, c["rects"])
    ref = ob.solve(nlp, c["x_init"], c["ref_states"], c["ref_inputs"])
    geo = ob.solve(nlp, c["x_init"], c["ref_states"], c["ref_inputs"], geometric_start=True)
    assert ref["status"] == 0 and geo["status"] == 0 and geo["iters"] < ref["iters"]
    assert np.abs(geo["states"] - c["states"]).max() <= Z_TOL and np.abs(geo["inputs"] - c["inputs"]).max() <= Z_TOL
    e_ref = emu.obca_solve_batch(cfg, obs, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    e_geo = emu.obca_solve_batch(cfg, obs_geo, c["x_init"][None], c["ref_states"][None], c["ref_inputs"][None])
    assert e_geo["status"][0] == 0 and e_geo["iters"][0] < e_ref["iters"][0]
    xs, us = split_z(e_geo["z"][0], cfg.horizon)
    assert np.abs(xs - geo["states"]).max() <= Z_TOL and np.abs(us - geo["inputs"]).max() <= Z_TOL
    assert abs(e_geo["obj"][0] - c["obj"]) <= OBJ_REL_TOL * abs(c["obj"])
