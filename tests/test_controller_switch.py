"""The caller-side controller switch of simulation.py:416-436,501-526 (collision test of the previous prediction ->
obstacle-aware or plain controller).  CPU only: the geometry is checked against the independent polygon-distance
routine of tests/geometry.py and against a literal per-state separating-axis loop; the switch with recording stand-ins
for the two controllers."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import geometry  # noqa: E402

from car_trailer_mpc_b200 import parking_lot_obstacles  # noqa: E402
from car_trailer_mpc_b200.collision import check_state_collision, check_trajectory_collision, stage_collisions  # noqa: E402
from car_trailer_mpc_b200.mpc_control_switch import SwitchingController  # noqa: E402

PARAMS = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05, "horizon": 50}


def _rects(obstacles):
    return [(o["center"][0], o["center"][1], o["width"], o["height"]) for o in obstacles]


def _sat_loop(state, obstacles):
    """One state, the slow way: corners of both bodies, 4 candidate axes per (body, obstacle), strict gap test."""
    veh, trl = geometry.body_corners(np.asarray(state, dtype=np.float64)[None, :])
    for body in (veh[0], trl[0]):
        e1, e2 = body[1] - body[0], body[3] - body[0]
        axes = [np.array([1.0, 0.0]), np.array([0.0, 1.0]), np.array([-e1[1], e1[0]]) / np.linalg.norm(e1),
                np.array([-e2[1], e2[0]]) / np.linalg.norm(e2)]
        for r in _rects(obstacles):
            box = geometry.box_corners(r)
            if not any((body @ a).max() < (box @ a).min() or (box @ a).max() < (body @ a).min() for a in axes):
                return True
    return False


def test_collision_flag_agrees_with_polygon_distance_and_sat_loop(traj):
    S, _ = traj
    obstacles = parking_lot_obstacles()
    rng = np.random.default_rng(11)
    # states along the shipped path pushed sideways: clear, grazing and overlapping poses all occur
    k = rng.integers(0, S.shape[1], size=600)
    X = S[:, k].copy()
    X[:2] += rng.normal(0.0, 1.5, size=(2, k.size))
    X[2:4] += rng.normal(0.0, 0.2, size=(2, k.size))
    flags = stage_collisions(X, PARAMS, obstacles)
    clear = geometry.clearance(X.T, _rects(obstacles))
    assert 50 < flags.sum() < 550                     # both outcomes are well represented
    decided = np.abs(clear) > 1e-9                    # exact contact is measure zero; skip it
    assert np.array_equal(flags[decided], clear[decided] < 0.0)
    for i in range(0, k.size, 7):
        assert check_state_collision(X[:, i], PARAMS, obstacles) == _sat_loop(X[:, i], obstacles)
    assert check_trajectory_collision(X, PARAMS, obstacles)
    assert not check_trajectory_collision(X[:, ~flags], PARAMS, obstacles)
    assert not check_trajectory_collision(X, PARAMS, [])          # simulation.py:375-376
    # touching without a gap counts as a collision (strict gap test, simulation.py:298): vehicle front edge on an obstacle edge
    exact = dict(PARAMS, L1=8.0)                                    # binary-exact lengths: front edge at x = 8 exactly
    wall = [{"center": (9.0, 0.0), "width": 2.0, "height": 10.0}]   # left edge at x = 8 exactly
    assert check_state_collision(np.array([0.0, 0.0, 0.0, 0.0]), exact, wall)
    assert not check_state_collision(np.array([-1e-9, 0.0, 0.0, 0.0]), exact, wall)


class _Recorder:
    def __init__(self, name, states):
        self.name, self.states, self.calls = name, states, 0

    def solve(self, x, xs, us):
        self.calls += 1
        return self.states.copy(), np.zeros((2, self.states.shape[1] - 1))


def test_switch_follows_the_previous_prediction(capsys):
    obstacles = [{"center": (30.0, 0.0), "width": 4.0, "height": 4.0}]
    N = 5
    clear = np.zeros((6, N + 1)); clear[0] = np.linspace(0.0, 1.0, N + 1)           # far from the obstacle
    hits = clear.copy(); hits[0] += 25.0                                            # vehicle front reaches x = 33
    obs_c, plain_c = _Recorder("obs", clear), _Recorder("plain", hits)
    sw = SwitchingController(obs_c, plain_c, PARAMS, obstacles)
    x0 = np.zeros(6); us = np.zeros((2, N))
    # step 1: no previous prediction -> the REFERENCE window is tested (simulation.py:503); it is clear -> plain controller
    sw.solve(x0, clear, us)
    assert (plain_c.calls, obs_c.calls) == (1, 0)
    # step 2: the plain controller's prediction collides -> obstacle-aware controller, and the reference prints
    sw.solve(x0, clear, us)
    assert (plain_c.calls, obs_c.calls) == (1, 1)
    assert "Using obstacle-aware MPC" in capsys.readouterr().out
    # step 3: the obstacle-aware prediction was clear -> back to the plain controller
    sw.solve(x0, clear, us)
    assert (plain_c.calls, obs_c.calls) == (2, 1)
    assert sw.used_obstacle_aware == [False, True, False]
    # a colliding reference window selects the obstacle-aware controller on the very first step
    sw.reset()
    sw.solve(x0, hits, us)
    assert obs_c.calls == 2 and sw.used_obstacle_aware == [True]
    # the stored prediction is a copy (the drivers reuse their buffers, simulation.py:463-464)
    assert sw.prev_mpc_prediction is not obs_c.states and np.array_equal(sw.prev_mpc_prediction, clear)


class _CoreController:
    """solve() through the host build of the kernel cores (tools/emu.py): plain tracking or obstacle-aware."""

    def __init__(self, cfg, obstacles=None):
        self.cfg, self.obstacles = cfg, obstacles
        self.last_status = self.last_iterations = 0

    def solve(self, x0, xs, us):
        import emu
        from car_trailer_mpc_b200 import problem as pb
        N = self.cfg.horizon
        xs_ = np.ascontiguousarray(np.asarray(xs).T)[None]
        us_ = np.ascontiguousarray(np.asarray(us).T)[None]
        x0_ = np.asarray(x0, dtype=np.float64)[None]
        r = (emu.solve_batch(self.cfg, x0_, xs_, us_) if self.obstacles is None
             else emu.obca_solve_batch(self.cfg, self.obstacles, x0_, xs_, us_))
        self.last_status, self.last_iterations = int(r["status"][0]), int(r["iters"][0])
        X, U = pb.unpack_z(r["z"], N)
        return X[0].T.copy(), U[0].T.copy()


def test_closed_loop_with_the_switch_clears_an_obstacle_the_plain_controller_hits(traj):
    """simulation.py's USE_SWITCH_MPC loop, headless: a 1 m box is put 0.1 m into the vehicle's swept path 25 steps
    ahead.  The plain controller drives through it; with the switch the obstacle-aware controller takes the steps whose
    previous prediction collides, and the executed path stays clear."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from car_trailer_mpc_b200 import closed_loop as cl, tracking_preset
    from car_trailer_mpc_b200.config import Obstacles
    S, U = traj
    N, k0 = 12, 40
    params = dict(PARAMS, horizon=N)
    x, y, th = S[k0 + 25, :3]
    front = np.array([x + np.cos(th) * 7.05 / 2, y + np.sin(th) * 7.05 / 2])
    side = np.array([-np.sin(th), np.cos(th)])
    c = front + side * (3.05 / 2 - 0.1 + 0.5 * (abs(np.cos(th)) + abs(np.sin(th))))
    rect = (c[0], c[1], 1.0, 1.0)
    obstacle_list = [{"center": (c[0], c[1]), "width": 1.0, "height": 1.0}]
    cfg = tracking_preset(N); cfg.max_iter = 300

    def run(controller):
        return cl.simulate_single(controller, S[k0:], U[k0:], S[k0], 2.5, 0.05, N, params)

    plain = run(_CoreController(cfg))
    sw = SwitchingController(_CoreController(cfg, Obstacles.from_list([rect])), _CoreController(cfg), params, obstacle_list,
                             verbose=False)
    switched = run(sw)
    assert plain.failures == 0 and switched.failures == 0
    assert geometry.clearance(plain.states, [rect]).min() < 0.0          # tracking alone collides
    assert geometry.clearance(switched.states, [rect]).min() > 0.0       # the switch keeps the executed path clear
    flags = np.array(sw.used_obstacle_aware)
    assert not flags[0] and flags.any() and not flags.all()
    # an obstacle-aware solve returns a clear prediction, so the reference's rule hands the next step back to the plain controller
    assert not (flags[1:] & flags[:-1]).any()
    # without obstacles the switch is the plain controller, bit for bit
    sw0 = SwitchingController(_CoreController(cfg, Obstacles.from_list([rect])), _CoreController(cfg), params, [], verbose=False)
    same = run(sw0)
    assert not any(sw0.used_obstacle_aware) and np.array_equal(same.states, plain.states)
