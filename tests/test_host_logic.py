"""Host-side logic around the solve (windows, layouts, index quirk, shift, interpolation, plant) against the
oracle's C restatements and direct loop restatements of the reference code."""
import math

import numpy as np

from car_trailer_mpc_b200 import problem as pb
from car_trailer_mpc_b200 import tracking_preset
from oracle import oracle


def test_interpolation_matches_loop_restatement(traj):
    import os
    S0 = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt"))
    U0 = np.loadtxt(os.path.join(pb.DATA_DIR, "input_traj.txt"))
    # literal loop of simulation.py:201-218
    N = U0.shape[1]; n = math.floor(0.1 / 0.05)
    Sn = np.zeros((6, n * N + 1)); Un = np.zeros((2, n * N))
    for k in range(N):
        for m in range(n):
            t = m / n
            Sn[:, k * n + m] = (1 - t) * S0[:, k] + t * S0[:, k + 1]
            Un[:, k * n + m] = U0[:, k]
    Sn[:, -1] = S0[:, -1]
    S, U = traj
    assert S.shape == (401, 6) and U.shape == (400, 2)
    assert np.array_equal(S, Sn.T) and np.array_equal(U, Un.T)


def test_window_regimes(traj):
    S, U = traj
    for N in (10, 40, 50):
        for k in (0, 1, 399 - N, 400 - N, 401 - N, 380, 399, 400, 401, 500):
            if k < 0:
                continue
            xs, us = pb.window(S, U, k, N)
            xo, uo = oracle.window(S, U, k, N)
            assert np.array_equal(xs, xo) and np.array_equal(us, uo), (N, k)
    xs, us = pb.window(S, U, 390, 40)           # tail: last state repeated, LAST INPUT repeated
    assert np.array_equal(xs[10:], np.repeat(S[400][None], 31, 0)) and np.array_equal(us[10:], np.repeat(U[399][None], 30, 0))
    xs, us = pb.window(S, U, 400, 40)           # past the end: ZERO input
    assert (us == 0).all() and np.array_equal(xs, np.repeat(S[400][None], 41, 0))
    ks = np.array([0, 5, 361, 390, 400, 450])
    xb, ub = pb.windows_batch(S, U, ks, 40)
    for i, k in enumerate(ks):
        a, b = pb.window(S, U, int(k), 40)
        assert np.array_equal(xb[i], a) and np.array_equal(ub[i], b)


def test_float_index_quirk():
    ks = pb.time_indices(40.0, 0.05)
    assert len(ks) == 801
    assert (ks != np.arange(801)).sum() == 199
    assert ks[6] == 5
    assert len(pb.time_indices(25.0, 0.05)) == 500


def test_pack_unpack_roundtrip_and_layout():
    rng = np.random.default_rng(3)
    N = 7
    X = rng.normal(size=(4, N + 1, 6)); U = rng.normal(size=(4, N, 2))
    z = pb.pack_z(X, U)
    assert z.shape == (4, 8 * N + 6)
    assert np.array_equal(z[0, :6], X[0, 0]) and np.array_equal(z[0, 6:8], U[0, 0]) and np.array_equal(z[0, 8:14], X[0, 1])
    assert np.array_equal(z[0, -6:], X[0, N])
    X2, U2 = pb.unpack_z(z, N)
    assert np.array_equal(X, X2) and np.array_equal(U, U2)


def test_shift_matches_reference_slicing():
    rng = np.random.default_rng(4)
    N = 5
    z = rng.normal(size=8 * N + 6)
    # literal restatement of mpc_control_nmpc.py:69-88 on a flat vector
    step = 8
    shifted = []
    for k in range(N - 1):
        shifted.extend(z[(k + 1) * step:(k + 1) * step + step])
    last_state = z[-step:-2]; last_input = z[-2:]
    shifted.extend(last_state); shifted.extend(last_input); shifted.extend(last_state)
    ref = np.array(shifted)
    assert ref.shape == z.shape
    assert np.array_equal(pb.shift_warm_start(z, N, reference_bug=True), ref)
    assert np.array_equal(oracle.shift(z, N, 1), ref)
    good = pb.shift_warm_start(z, N, reference_bug=False)
    assert np.array_equal(good, oracle.shift(z, N, 0))
    assert np.array_equal(good[-6:], z[-6:]) and np.array_equal(good[8 * (N - 1):8 * (N - 1) + 6], z[-6:])
    assert np.array_equal(good[8 * (N - 1) + 6:8 * N], z[8 * (N - 1) + 6:8 * N])


def test_se2_transform_keeps_dynamics_exact(traj):
    S, U = traj
    cfg = tracking_preset(40)
    Sm = pb.se2_transform(S, (10.0, -3.0, 1.0))
    d0 = pb.dynamics_defect(cfg, S, U); d1 = pb.dynamics_defect(cfg, Sm, U)
    # (the up-sampled trajectory is not Euler-exact at dt=0.05; the defect must merely be unchanged)
    c, s = math.cos(1.0 - S[0, 2]), math.sin(1.0 - S[0, 2])
    rot = np.stack([c * d0[:, 0] - s * d0[:, 1], s * d0[:, 0] + c * d0[:, 1]], 1)
    assert np.abs(d1[:, :2] - rot).max() < 1e-12 and np.abs(d1[:, 2:] - d0[:, 2:]).max() < 1e-12
    assert abs(Sm[0, 0] - 10.0) < 1e-12 and abs(Sm[0, 2] - 1.0) < 1e-12


def test_scenarios_are_deterministic_and_feasible():
    cfg = tracking_preset(40)
    a = pb.make_scenarios(cfg, 100); b = pb.make_scenarios(cfg, 100)
    assert np.array_equal(a.x_init, b.x_init) and np.array_equal(a.ref_states, b.ref_states)
    lb, ub = np.array(cfg.x_lb[:]), np.array(cfg.x_ub[:])
    assert (a.x_init[:, 2:] > lb[2:]).all() and (a.x_init[:, 2:] < ub[2:]).all()
    assert set(np.unique(a.family)) <= {-1, 0, 1, 2, 3, 4, 5, 6}


def test_plant_step_oracle_matches_reference_formulas():
    cfg = tracking_preset(40)
    rng = np.random.default_rng(5)
    q = np.array([1.0, 2.0, 0.3, 0.2, -0.3, -2.0]); u = np.array([0.5, -0.2])
    # nominal: plain Euler (simulation.py:187-194 with disturbance_params=None)
    f = pb.model_f(cfg, q, u)
    assert np.allclose(oracle.plant_step(cfg, q, u), q + f * cfg.dt, rtol=0, atol=1e-15)
    # disturbed (simulation.py:26-32 defaults): friction/slippage scale u, lateral drift along theta+pi/2
    d = [0.9, 0.9, 0.01, 0.0]
    fd = pb.model_f(cfg, q, u * 0.9)
    exp = q + fd * cfg.dt
    mag = 0.01 * abs(q[5]) * abs(q[4])
    exp[0] += mag * math.cos(q[2] + math.pi / 2) * cfg.dt
    exp[1] += mag * math.sin(q[2] + math.pi / 2) * cfg.dt
    assert np.allclose(oracle.plant_step(cfg, q, u, d), exp, rtol=0, atol=1e-15)
    n = rng.normal(size=6)
    assert np.allclose(oracle.plant_step(cfg, q, u, d, n, cfg.dt), exp + n * cfg.dt, rtol=0, atol=1e-15)
