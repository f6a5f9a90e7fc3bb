"""Closed-loop parity on the GPU: identical jackknife / no-jackknife outcomes and final errors between the CUDA
path (shim classes and the batched on-device driver) and the CPU oracle over full simulation.py-style runs."""
import math

import numpy as np
import pytest

from test_closed_loop_cpu import PARAMS, OracleController

from car_trailer_mpc_b200 import closed_loop as cl
from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200 import tracking_preset

pytestmark = pytest.mark.gpu

SB = {"lb": [-np.inf, -np.inf, -np.pi, -np.pi / 3, -np.pi / 4, -10.0], "ub": [np.inf, np.inf, np.pi, np.pi / 3, np.pi / 4, 10.0]}
IB = {"lb": [-5, -np.pi / 2], "ub": [5, np.pi / 2]}


def test_full_simulation_run_shim_vs_oracle(traj):
    """Config 1: B=1, N=50, T_sim=40 s (801 steps), nominal plant, start on the trajectory."""
    import torch
    assert torch.cuda.is_available()
    from car_trailer_mpc_b200 import MPCTrackingControl, TruckTrailerModel
    S, U = traj
    N = 50
    params = dict(PARAMS, horizon=N)
    ctl = MPCTrackingControl(TruckTrailerModel(params), params, np.eye(6), 10 * np.eye(2), SB, IB)
    gpu = cl.simulate_single(ctl, S, U, S[0], 40.0, 0.05, N, params)
    ref = cl.simulate_single(OracleController(tracking_preset(N)), S, U, S[0], 40.0, 0.05, N, params)
    mg, mr = gpu.metrics(S[-1]), ref.metrics(S[-1])
    assert mg["jackknife"] == mr["jackknife"] == False and mg["failures"] == mr["failures"] == 0
    assert np.abs(gpu.controls - ref.controls).max() <= 1e-4          # every applied control, all 801 steps
    assert np.abs(gpu.states - ref.states).max() <= 1e-5
    for key in ("distance_error", "heading_error", "hitch_error", "max_abs_psi"):
        assert abs(mg[key] - mr[key]) <= 1e-5, key
    assert abs(mg["distance_error"] - 0.0720) < 2e-4 and abs(mg["max_abs_psi"] - 0.7488) < 2e-4
    assert gpu.iterations == ref.iterations


def test_batched_closed_loop_vs_single_episodes(traj):
    """The on-device batch driver (shared-trajectory windows in-kernel, plant kernel, counter-based noise) against
    per-scenario host loops with the oracle, nominal and disturbed, incl. a start that jackknifes."""
    import torch
    from car_trailer_mpc_b200 import BatchSolver
    S, U = traj
    N = 40
    cfg = tracking_preset(N)
    params = dict(PARAMS, horizon=N)
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    B = 6
    x0 = S[0][None] + rng.normal(0, 0.05, size=(B, 6)) * np.array([1, 1, 0.2, 0.2, 0.1, 1])
    solver = BatchSolver(cfg, 0)
    T_sim = 6.0
    for dist in (None, cl.DEFAULT_DISTURBANCE):
        out = cl.simulate_batch(solver, S, U, torch.from_numpy(x0).to(dev), T_sim, 0.05, dist, seed=11, record_every=1)
        hist = out["history"].cpu().numpy()          # [steps, B, 6]
        ks = pb.time_indices(T_sim, 0.05)
        for b in range(B):
            # host replica with the same counter-based measurement noise
            ctl = OracleController(cfg)
            host_fail = 0
            state = x0[b].copy()
            for step, k in enumerate(ks):
                xs, us = pb.window(S, U, int(k), N)
                meas = state
                if dist is not None:
                    n = cl.counter_normal(11, 2 * step, torch.tensor([b]), 6)[0].numpy()
                    meas = state + dist["process_noise_std"] * n
                r = ctl.solve(meas, xs.T, us.T)
                host_fail += ctl.last_status > 1
                state = cl.plant_update(state, r[1][:, 0], params, dist)
                assert np.abs(state - hist[step, b]).max() <= 1e-5, (dist is not None, b, step)
            assert bool(out["jackknife"][b]) == bool(np.abs(hist[:, b, 3]).max() > cl.JACKKNIFE_LIMIT)
            assert int(out["failures"][b]) == host_fail
    assert out["steps"] == len(ks)


def test_device_episode_kernel_vs_host_driven_loop(traj):
    """ttmpc_episode_batch (whole closed loop inside one persistent kernel, in-kernel counter-based noise) against the
    host-driven batched loop (one solve + plant launch per step, torch noise): same metrics for every scenario."""
    import torch
    from car_trailer_mpc_b200 import BatchSolver
    S, U = traj
    cfg = tracking_preset(40); cfg.max_iter = 100
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(9)
    B = 200   # more scenarios than one warp, ragged
    x0 = S[0][None] + rng.normal(0, 0.03, size=(B, 6)) * np.array([1, 1, 0.2, 0.2, 0.1, 1])
    x0[7, 4] = 0.9    # steering angle beyond its bound, standing still: every solve fails, zero control keeps it there ->
    x0[7, 5] = 0.0    # the nmpc variant stops this run after 21 consecutive failures (simulation_nmpc.py:212-216)
    x0_d = torch.from_numpy(x0).to(dev)
    ids = torch.arange(1000, 1000 + B, device=dev, dtype=torch.int64)
    solver = BatchSolver(cfg, 0)
    T_sim = 4.0
    ks = pb.time_indices(T_sim, 0.05)
    for dist, variant in ((None, "tracking"), (cl.DEFAULT_DISTURBANCE, "tracking"), (cl.DEFAULT_DISTURBANCE, "nmpc")):
        host = cl.simulate_batch(solver, S, U, x0_d, T_sim, 0.05, dist, seed=21, scenario_ids=ids, variant=variant)
        devr = solver.episodes(x0_d, S, U, ks, dist, variant=variant, seed=21, scenario_ids=ids)
        torch.cuda.synchronize()
        m = devr["metrics"].cpu().numpy()
        assert np.abs(devr["final_state"].cpu().numpy() - host["final_state"].cpu().numpy()).max() < 1e-6
        assert np.abs(m[:, 0] - host["distance_error"].cpu().numpy()).max() < 1e-6
        assert np.abs(m[:, 1] - host["heading_error"].abs().cpu().numpy()).max() < 1e-6
        assert np.abs(m[:, 2] - host["hitch_error"].abs().cpu().numpy()).max() < 1e-6
        assert np.abs(m[:, 3] - host["max_abs_psi"].cpu().numpy()).max() < 1e-6
        assert np.array_equal(m[:, 4] > 0.5, host["jackknife"].cpu().numpy())
        assert np.array_equal(m[:, 5].astype(int), host["failures"].cpu().numpy())
        assert np.abs(m[:, 6] - host["mean_iters"].cpu().numpy()).max() < 0.05
        assert np.abs(m[:, 7] - host["rms_tracking_error"].cpu().numpy()).max() < 1e-6
        if variant == "nmpc":
            assert int(m[7, 5]) == 21 and int(host["failures"][7]) == 21     # stopped, not 81 failed steps


def test_nmpc_shim_closed_loop_vs_oracle(traj):
    """simulation_nmpc.py run (N=30, 500 steps, warm start with the reference's slicing) through the TruckTrailerNMPC
    shim vs the oracle under the same driver: same outcome.  At tol 1e-3 both stop at the same iterate because they run
    the same algorithm; a flipped termination would show up as a 1e-3-level control difference."""
    from test_closed_loop_cpu import OracleNMPC
    from car_trailer_mpc_b200 import TruckTrailerModel, TruckTrailerNMPC, nmpc_preset
    S, U = traj
    N = 30
    params = dict(PARAMS, horizon=N)
    Qn, Rn = np.diag([1.0, 1.0, 2.0, 3.0, 1.0, 1.0]), np.diag([5.0, 8.0])
    sbn = {"lb": [-np.inf, -np.inf, -np.pi, -np.pi / 3, -np.pi / 4, -8.0], "ub": [np.inf, np.inf, np.pi, np.pi / 3, np.pi / 4, 8.0]}
    ibn = {"lb": [-4, -np.pi / 2], "ub": [4, np.pi / 2]}
    for dist in (None, cl.DEFAULT_DISTURBANCE):
        ctl = TruckTrailerNMPC(TruckTrailerModel(params), params, Qn, Rn, sbn, ibn, shift_reference_bug=True)
        gpu = cl.simulate_single(ctl, S, U, S[0], 25.0, 0.05, N, params, dist, np.random.RandomState(5), variant="nmpc")
        ref = cl.simulate_single(OracleNMPC(nmpc_preset(N), True), S, U, S[0], 25.0, 0.05, N, params, dist,
                                 np.random.RandomState(5), variant="nmpc")
        mg, mr = gpu.metrics(S[-1]), ref.metrics(S[-1])
        assert mg["jackknife"] == mr["jackknife"] and mg["failures"] == mr["failures"] and mg["steps"] == mr["steps"] == 500
        assert np.abs(gpu.controls - ref.controls).max() < 1e-4
        assert np.abs(gpu.states - ref.states).max() < 1e-4


def test_fuzzy_controller_shim_vs_oracle(traj):
    """simulation_fuzzy.py configuration (N=40, Q=I, R=10I, tol 1e-3, warm start, fuzzy weight scalings as per-problem
    kernel inputs) through the MPCTrackingControlFuzzy shim vs the oracle with explicitly scaled Q/R matrices."""
    import torch
    from car_trailer_mpc_b200 import MPCTrackingControlFuzzy, TruckTrailerModel, nmpc_preset
    from car_trailer_mpc_b200.mpc_control_fuzzy import fuzzy_weights
    from oracle import oracle
    assert torch.cuda.is_available()
    S, U = traj
    N = 40
    params = dict(PARAMS, horizon=N)
    ctl = MPCTrackingControlFuzzy(TruckTrailerModel(params), params, np.eye(6), 10 * np.eye(2), SB, IB)
    base = nmpc_preset(N)
    base.set_weights(np.eye(6), 10 * np.eye(2))
    base.set_bounds(SB["lb"], SB["ub"], IB["lb"], IB["ub"])

    class OracleFuzzy:
        def __init__(self):
            self._last = None
            self.last_status = self.last_iterations = None

        def solve(self, x, ref_s, ref_u):
            q, r = fuzzy_weights(x, ref_s)
            zw = None if self._last is None else pb.shift_warm_start(self._last, N, True)
            for qq, rr in ((q, r), (np.ones(6), np.ones(2))):
                c = base.copy()
                c.set_weights(np.diag(qq) @ np.eye(6) @ np.diag(qq), np.diag(rr) @ (10 * np.eye(2)) @ np.diag(rr))
                res = oracle.solve(c, np.asarray(x), np.ascontiguousarray(ref_s.T), np.ascontiguousarray(ref_u.T), z_warm=zw)
                self.last_status, self.last_iterations = int(res["status"]), int(res["iters"])
                if self.last_status <= 1:
                    break
            if self.last_status > 1:
                return None, None
            self._last = res["z"]
            X, Uu = pb.unpack_z(res["z"], N)
            return X.T.copy(), Uu.T.copy()

    gpu = cl.simulate_single(ctl, S, U, S[0], 30.0, 0.05, N, params, variant="nmpc")
    ref = cl.simulate_single(OracleFuzzy(), S, U, S[0], 30.0, 0.05, N, params, variant="nmpc")
    mg, mr = gpu.metrics(S[-1]), ref.metrics(S[-1])
    assert mg["steps"] == mr["steps"] == 600 and mg["failures"] == mr["failures"] == 0
    assert mg["jackknife"] == mr["jackknife"] == False
    assert np.abs(gpu.controls - ref.controls).max() < 1e-4 and np.abs(gpu.states - ref.states).max() < 1e-4
    q, r = ctl.last_weights
    assert q.max() > 1.0 and r.max() > 1.0    # the rule base was active (hitch angle / reversing)


def test_weighted_batch_solve_device_and_host_paths():
    import torch
    from car_trailer_mpc_b200 import BatchSolver
    from car_trailer_mpc_b200.mpc_control_fuzzy import fuzzy_weights
    from oracle import oracle
    cfg = tracking_preset(40)
    sc = pb.make_scenarios(cfg, 300, seed=29)
    qw = np.empty((300, 6)); rw = np.empty((300, 2))
    for i in range(300):
        qw[i], rw[i] = fuzzy_weights(sc.x_init[i], sc.ref_states[i].T)
    s = BatchSolver(cfg, 0)
    dev = torch.device("cuda:0")
    a = s.solve(sc.x_init, sc.ref_states, sc.ref_inputs, q_weights=qw, r_weights=rw)
    t = lambda x: torch.from_numpy(x).to(dev)
    b = s.solve(t(sc.x_init), t(sc.ref_states), t(sc.ref_inputs), q_weights=t(qw), r_weights=t(rw))
    torch.cuda.synchronize()
    assert np.array_equal(a["z"], b["z"].cpu().numpy()) and (a["status"] == 0).all()
    for i in range(0, 300, 17):
        c = cfg.copy()
        c.set_weights(np.diag(qw[i] ** 2), 10 * np.diag(rw[i] ** 2))
        ref = oracle.solve(c, sc.x_init[i], sc.ref_states[i], sc.ref_inputs[i])
        assert np.abs(ref["u0"] - a["u0"][i]).max() <= 1e-4 and abs(ref["obj"] - a["obj"][i]) <= 1e-6 * max(1.0, abs(ref["obj"]))
    unweighted = s.solve(sc.x_init, sc.ref_states, sc.ref_inputs)
    assert np.abs(unweighted["u0"] - a["u0"]).max() > 1e-3      # the scalings change the solution
    dense = tracking_preset(40); Q = np.eye(6); Q[0, 1] = Q[1, 0] = 0.2; dense.set_weights(Q, 10 * np.eye(2))
    with pytest.raises(Exception):
        BatchSolver(dense, 0).solve(sc.x_init[:4], sc.ref_states[:4], sc.ref_inputs[:4], q_weights=qw[:4], r_weights=rw[:4])
