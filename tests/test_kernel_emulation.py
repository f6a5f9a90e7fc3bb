"""The device solver core (csrc/ttmpc_core.cuh: fused backward sweep with costate recursion, structured
Riccati, forward and trial sweeps), compiled for the host by tools/kernel_emu.cpp, must agree with the
oracle -- two independently written implementations of the same algorithm.  Runs without a GPU."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emu  # noqa: E402
from parity import assert_parity  # noqa: E402

from car_trailer_mpc_b200 import nmpc_preset, tracking_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402
from oracle import oracle  # noqa: E402


def test_emulated_kernel_matches_oracle_narrow_and_wide():
    cfg = tracking_preset(40); cfg.max_iter = 200
    for sig in (pb.SIGMA_NARROW, pb.SIGMA_WIDE):
        sc = pb.make_scenarios(cfg, 256, sigma=sig)
        r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, nthreads=4)
        r1 = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
        assert_parity(cfg, r1, r0, sc.x_init)
        assert (r0["iters"] == r1["iters"]).mean() > 0.98


def test_emulated_kernel_other_horizons_and_presets():
    for cfg in (tracking_preset(10), tracking_preset(73), nmpc_preset(30)):
        sc = pb.make_scenarios(cfg, 48, seed=5)
        r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, nthreads=4)
        r1 = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
        if cfg.tol < 1e-6:
            assert_parity(cfg, r1, r0, sc.x_init)
        else:  # tol 1e-3: both stop within ~1e-3 of the KKT point; compare loosely
            assert np.array_equal(r0["status"], r1["status"])
            assert np.abs(r0["u0"] - r1["u0"]).max() < 1e-6


def test_emulated_kernel_warm_start_and_general_weights():
    cfg = nmpc_preset(30)
    Q = np.diag([1.0, 1.0, 2.0, 3.0, 1.0, 1.0]); Q[0, 1] = Q[1, 0] = 0.3; Q[2, 5] = Q[5, 2] = -0.2
    R = np.array([[5.0, 0.7], [0.7, 8.0]])
    cfg.set_weights(Q, R); cfg.tol = 1e-8; cfg.acceptable_tol = 1e-6; cfg.acceptable_iter = 15
    sc = pb.make_scenarios(cfg, 32, seed=9, families=False)
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    zw = pb.shift_warm_start(r0["z"], 30, reference_bug=True)
    r0w = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, z_warm=zw)
    r1w = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, z_warm=zw)
    assert_parity(cfg, r1w, r0w, sc.x_init)
    assert_parity(cfg, r1w, r0, sc.x_init)   # warm and cold starts reach the same minimiser


def test_infeasible_x0_policy():
    cfg = tracking_preset(20)
    sc = pb.make_scenarios(cfg, 4, seed=2, families=False)
    sc.x_init[1, 4] = 0.9   # phi beyond pi/4: reference NLP infeasible
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    r1 = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    assert r0["status"][1] == 5 and r1["status"][1] == 5
    assert np.array_equal(r0["status"], r1["status"])


def test_generic_bound_code_path_equals_specialised_one():
    cfg = tracking_preset(40)
    sc = pb.make_scenarios(cfg, 64, seed=13, sigma=pb.SIGMA_WIDE)
    a = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    b = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, force_generic=True)
    for key in ("z", "u0", "obj", "iters", "status"):
        assert np.array_equal(a[key], b[key]), key
    # general masks: x,y bounded, theta free, v one-sided
    cfg.set_bounds([-200.0, -200.0, -np.inf, -1.0, -0.7, -np.inf], [200.0, 200.0, np.inf, 1.0, 0.7, 6.0], [-4.0, -1.0], [4.0, 1.0])
    sc = pb.make_scenarios(cfg, 64, seed=14, families=False)
    r0 = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    r1 = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    assert_parity(cfg, r1, r0, sc.x_init)


def test_shared_trajectory_mode_in_emulation(traj):
    S, U = traj
    cfg = tracking_preset(40)
    rng = np.random.default_rng(5)
    k = np.concatenate([rng.integers(0, 460, size=40), [0, 360, 361, 399, 400, 401, 1000]]).astype(np.int32)
    xs, us = pb.windows_batch(S, U, k, 40)
    x = xs[:, 0, :] + rng.normal(0, 0.02, size=(len(k), 6))
    a = emu.solve_batch(cfg, x, xs, us)
    b = emu.solve_batch(cfg, x, k_index=k, traj_states=S, traj_inputs=U)
    for key in ("z", "u0", "obj", "iters", "status"):
        assert np.array_equal(a[key], b[key]), key


def test_round_robin_line_search_equals_loop_line_search():
    """ipm_step_rr (one trial per round, episode kernel) must walk exactly the same iterates as ipm_step."""
    cfg = tracking_preset(30); cfg.max_iter = 60
    sc = pb.make_scenarios(cfg, 96, seed=17, sigma=pb.SIGMA_WIDE * 3.0)   # far off: backtracking and failures occur
    a = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    b = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, force_generic=4)
    for key in ("z", "u0", "obj", "iters", "status"):
        assert np.array_equal(a[key], b[key], equal_nan=True), key
    assert (a["iters"] > 12).any()


def test_speculative_first_trial_walks_the_classic_iterates():
    """ipm_step hands the first trial point of a line search to the next backward sweep untested and the test is made
    from that sweep's statistics (a rejection restarts the problem with classic trial sweeps -- modes 1, 2 -- or resumes
    the classic search from the intact previous iterate -- mode 3, ping-pong copies): results must be those of
    the classic trial-sweep line search bit for bit -- well-posed batches, warm starts, and a far-off batch with
    backtracking, line-search failures and infeasible x_0.  The classic side is the core as the shipped library compiles
    it; the speculative side is the experiment configuration -DTTMPC_SPECULATION=1 (force_generic bits 3-4 =
    Params::speculate, a second host library built by tools/emu.py)."""
    cfg = tracking_preset(40); cfg.max_iter = 200
    sc = pb.make_scenarios(cfg, 192, seed=5, sigma=pb.SIGMA_WIDE)
    rng = np.random.default_rng(3)
    far = sc.x_init + rng.normal(0, 1, size=sc.x_init.shape) * np.array([2, 2, 0.5, 0.6, 0.5, 3])
    cold = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    warm = pb.shift_warm_start(cold["z"], 40, reference_bug=True)
    seen = set()
    for x, zw in ((sc.x_init, None), (sc.x_init, warm), (far, None)):
        a = emu.solve_batch(cfg, x, sc.ref_states, sc.ref_inputs, z_warm=zw)
        for mode in (1, 2, 3):   # 3: every step speculated into the second copy of the iterate rows, rejections resumed
            b = emu.solve_batch(cfg, x, sc.ref_states, sc.ref_inputs, z_warm=zw, force_generic=8 * mode)
            for key in ("z", "u0", "obj", "iters", "status"):
                assert np.array_equal(a[key], b[key], equal_nan=True), (mode, key)
        seen |= set(a["status"].tolist())
    assert {0, 3, 5} <= seen   # converged, line-search failure and infeasible-x0 instances were all exercised


def test_per_problem_weights_match_oracle_with_scaled_matrices():
    """PW kernels (mpc_control_fuzzy.py's parametric weights): Q_w = diag(q) Q diag(q), R_w = diag(r) R diag(r)."""
    from car_trailer_mpc_b200.mpc_control_fuzzy import fuzzy_weights
    cfg = tracking_preset(30)
    cfg.set_weights(np.diag([1.0, 1.0, 2.0, 3.0, 1.0, 1.0]), np.diag([5.0, 8.0]))
    sc = pb.make_scenarios(cfg, 24, seed=23, families=False)
    qw = np.empty((24, 6)); rw = np.empty((24, 2))
    for i in range(24):
        qw[i], rw[i] = fuzzy_weights(sc.x_init[i], sc.ref_states[i].T)
    qw[0], rw[0] = 1.0, 1.0
    got = emu.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, q_weights=qw, r_weights=rw)
    assert (qw > 1.0).any() and (rw > 1.0).any()
    for i in range(24):
        c = cfg.copy()
        c.set_weights(np.diag(qw[i]) @ cfg.Qm() @ np.diag(qw[i]), np.diag(rw[i]) @ cfg.Rm() @ np.diag(rw[i]))
        ref = oracle.solve(c, sc.x_init[i], sc.ref_states[i], sc.ref_inputs[i])
        assert ref["status"] == got["status"][i] == 0
        assert np.abs(ref["u0"] - got["u0"][i]).max() < 1e-7 and abs(ref["obj"] - got["obj"][i]) < 1e-9 * max(1, abs(ref["obj"]))
    plain = emu.solve_batch(cfg, sc.x_init[:1], sc.ref_states[:1], sc.ref_inputs[:1])
    assert np.array_equal(plain["z"][0], got["z"][0])   # unit scalings == unweighted kernel


def test_core_under_address_sanitizer(tmp_path):
    """compute-sanitizer is closed on the GPU pool, so the index arithmetic of the solver core (slot layout, carried
    storage, every kernel variant) is exercised on the host under ASan + UBSan instead."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "emu_asan")
    exe_spec = str(tmp_path / "emu_asan_spec")   # experiment configuration (two-copy scratch layout, speculative modes)
    for out_, extra in ((exe, []), (exe_spec, ["-DTTMPC_SPECULATION=1"])):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-DTTMPC_BANK=32", *extra, "-fsanitize=address,undefined",
                               "-fno-sanitize-recover=all", "-o", out_, os.path.join(root, "tools", "emu_asan_main.cpp"),
                               os.path.join(root, "tools", "kernel_emu.cpp"), "-lm"])
    cfg = tracking_preset(12)
    B = 35
    sc = pb.make_scenarios(cfg, B, seed=41, sigma=pb.SIGMA_WIDE)
    import ctypes
    open(tmp_path / "cfg.bin", "wb").write(bytes(ctypes.string_at(ctypes.byref(cfg), ctypes.sizeof(cfg))))
    rng = np.random.default_rng(2)
    for name, arr in (("x", sc.x_init), ("xs", sc.ref_states), ("us", sc.ref_inputs), ("qw", 1 + rng.random((B, 6))), ("rw", 1 + rng.random((B, 2)))):
        np.ascontiguousarray(arr, dtype=np.float64).tofile(tmp_path / f"{name}.bin")
    for cmd in ([exe, str(tmp_path), "12", str(B)], [exe_spec, str(tmp_path), "12", str(B), "spec"]):
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "checksum" in out.stdout and "nan" not in out.stdout.lower()
