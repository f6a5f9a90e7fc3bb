"""Closed-loop driver logic on CPU with an oracle-backed controller (test double with the reference's `solve`
interface): reproduces the survey's expected-outcome table for the nominal run from S[:,0],
the RNG draw order and the NMPC failure policy."""
import numpy as np

from car_trailer_mpc_b200 import closed_loop as cl
from car_trailer_mpc_b200 import nmpc_preset, tracking_preset
from car_trailer_mpc_b200 import problem as pb
from oracle import oracle

PARAMS = {"M": 0.15, "L1": 7.05, "L2": 12.45, "W1": 3.05, "W2": 2.95, "dt": 0.05}


class OracleController:
    """The CPU oracle behind the reference's controller interface (states [6,N+1], inputs [2,N])."""

    def __init__(self, cfg, none_on_failure=False):
        self.cfg, self.none_on_failure = cfg, none_on_failure
        self.last_status = self.last_iterations = None

    def solve(self, x, ref_s, ref_u):
        r = oracle.solve(self.cfg, np.asarray(x), np.ascontiguousarray(ref_s.T), np.ascontiguousarray(ref_u.T))
        self.last_status, self.last_iterations = int(r["status"]), int(r["iters"])
        if self.none_on_failure and self.last_status > 1:
            return None, None
        X, U = pb.unpack_z(r["z"], self.cfg.horizon)
        return X.T.copy(), U.T.copy()


def test_nominal_closed_loop_matches_survey_probe(traj):
    S, U = traj
    expect = {
        50: dict(max_psi=0.7488, dist=0.0720, head=0.0199, hitch=0.1340, u0=(-4.968355, 1.499396)),
        40: dict(max_psi=0.7515, dist=0.0995, head=0.0170, hitch=0.1371, u0=(-4.975572, 1.497000)),
    }
    for N, e in expect.items():
        cfg = tracking_preset(N)
        ep = cl.simulate_single(OracleController(cfg), S, U, S[0], 40.0, 0.05, N, dict(PARAMS, horizon=N))
        m = ep.metrics(S[-1])
        assert m["steps"] == 801 and m["failures"] == 0 and not m["jackknife"]
        assert abs(m["max_abs_psi"] - e["max_psi"]) < 2e-4
        assert abs(m["distance_error"] - e["dist"]) < 2e-4
        assert abs(m["heading_error"] - e["head"]) < 2e-4 and abs(m["hitch_error"] - e["hitch"]) < 2e-4
        assert np.abs(ep.controls[0] - np.array(e["u0"])).max() < 2e-6
        assert abs(m["max_abs_phi"] - 0.7854) < 1e-4            # phi bound active (steps 6-10)
        assert 5.0 <= np.mean(ep.iterations) <= 6.0 and max(ep.iterations) <= 16


def test_rng_order_and_disturbed_run_is_reproducible(traj):
    S, U = traj
    N = 30
    cfg = tracking_preset(N)
    a = cl.simulate_single(OracleController(cfg), S, U, S[0], 2.0, 0.05, N, dict(PARAMS, horizon=N),
                           cl.DEFAULT_DISTURBANCE, np.random.RandomState(3))
    b = cl.simulate_single(OracleController(cfg), S, U, S[0], 2.0, 0.05, N, dict(PARAMS, horizon=N),
                           cl.DEFAULT_DISTURBANCE, np.random.RandomState(3))
    assert np.array_equal(a.states, b.states)
    # 12 normals per step: 6 measurement + 6 discarded plant draws
    rs = np.random.RandomState(3)
    for _ in range(len(a.controls)):
        rs.normal(0, 0.02, 6); rs.normal(0, 0.02, 6)
    probe = np.random.RandomState(3)
    probe.normal(0, 0.02, 12 * len(a.controls))
    assert rs.normal() == probe.normal()
    assert not np.array_equal(a.states, cl.simulate_single(OracleController(cfg), S, U, S[0], 2.0, 0.05, N,
                                                           dict(PARAMS, horizon=N)).states)


def test_nmpc_failure_policy(traj):
    S, U = traj
    N = 30
    cfg = nmpc_preset(N)
    bad0 = S[0].copy(); bad0[4] = 1.2           # phi outside its bound: every solve fails
    ep = cl.simulate_single(OracleController(cfg, none_on_failure=True), S, U, bad0, 25.0, 0.05, N, dict(PARAMS, horizon=N),
                            variant="nmpc")
    assert ep.aborted and ep.failures == 21 and (ep.controls == 0).all()   # zero control, abort after >20 consecutive


def test_lqr_score_matches_definition(traj):
    S, U = traj
    cfg = tracking_preset(50)
    Q, R = np.eye(6), 10 * np.eye(2)
    xg = np.array([10.0, 5.0, 0.3, 0.1, 0.05, -1.0]); ug = np.zeros(2)   # a controllable linearisation point
    P = cl.lqr_riccati(cfg, Q, R, xg, ug)
    A, B = cl.euler_jacobians(cfg, xg, ug)
    res = A.T @ P @ A - P - A.T @ P @ B @ np.linalg.solve(R + B.T @ P @ B, B.T @ P @ A) + Q
    assert np.abs(res).max() < 1e-6 * np.abs(P).max()
    dx = np.array([0.1, -0.2, 0.01, 0.02, 0.0, 0.1])
    assert abs(cl.lqr_distance(xg + dx, xg, P) - dx @ P @ dx) < 1e-12
    # Jacobians agree with finite differences of the Euler map
    f = lambda x, u: x + cfg.dt * pb.model_f(cfg, x, u)
    h = 1e-6
    Afd = np.stack([(f(xg + h * e, ug) - f(xg - h * e, ug)) / (2 * h) for e in np.eye(6)], 1)
    assert np.abs(A - Afd).max() < 1e-8


class OracleNMPC(OracleController):
    """TruckTrailerNMPC semantics on the oracle: shifted warm start (reference slicing by default), (None, None) on
    failure without touching the stored solution (mpc_control_nmpc.py:90-113)."""

    def __init__(self, cfg, reference_bug=True):
        super().__init__(cfg, none_on_failure=True)
        self._last, self.bug = None, reference_bug

    def solve(self, x, ref_s, ref_u):
        zw = None if self._last is None else pb.shift_warm_start(self._last, self.cfg.horizon, self.bug)
        r = oracle.solve(self.cfg, np.asarray(x), np.ascontiguousarray(ref_s.T), np.ascontiguousarray(ref_u.T), z_warm=zw)
        self.last_status, self.last_iterations = int(r["status"]), int(r["iters"])
        if self.last_status > 1:
            return None, None
        self._last = r["z"]
        X, U = pb.unpack_z(r["z"], self.cfg.horizon)
        return X.T.copy(), U.T.copy()


def test_nmpc_closed_loop_runs_clean(traj):
    """simulation_nmpc.py configuration (N=30, T_sim=25 s, Q=diag(1,1,2,3,1,1), R=diag(5,8), tol 1e-3, warm start)."""
    S, U = traj
    N = 30
    cfg = nmpc_preset(N)
    for bug in (True, False):
        ep = cl.simulate_single(OracleNMPC(cfg, bug), S, U, S[0], 25.0, 0.05, N, dict(PARAMS, horizon=N), variant="nmpc")
        m = ep.metrics(S[-1])
        assert m["steps"] == 500 and m["failures"] == 0 and not m["jackknife"] and not ep.aborted
        assert m["distance_error"] < 0.2 and m["max_abs_psi"] < 0.75
