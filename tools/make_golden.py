"""Generates tests/golden/*.npz: seeded inputs + the oracle's outputs (the reference itself cannot be run
here -- CasADi/Ipopt are not installable -- so the golden solves are ORACLE outputs; the three
known-answer problems of SURVEY.md Appendix E.2 are included and were cross-checked with SciPy SLSQP,
tests/test_oracle_cross_solver.py).  Re-run with:  python tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from car_trailer_mpc_b200 import nmpc_preset, tracking_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402
from oracle import oracle  # noqa: E402

out_dir = os.path.join(ROOT, "tests", "golden")
os.makedirs(out_dir, exist_ok=True)
S, U = pb.load_reference_trajectory()


def save(name, cfg, x, xs, us, r, **extra):
    np.savez_compressed(os.path.join(out_dir, name), horizon=cfg.horizon, x_init=x, ref_states=xs, ref_inputs=us,
                        z=r["z"], u0=r["u0"], obj=r["obj"], kkt=r["kkt"], iters=r["iters"], status=r["status"], **extra)
    print(name, "status", np.bincount(r["status"], minlength=6), "iters mean", r["iters"].mean())


# 1. known answers (Appendix E.2) + 61 seeded scenarios, N=40, tracking preset
cfg = tracking_preset(40)
rng = np.random.default_rng(0)
xk, xsk, usk = [], [], []
for k0 in (0, 100, 200):
    xk.append(S[k0] + rng.normal(0, 0.02, 6))
    a, b = pb.window(S, U, k0, 40)
    xsk.append(a); usk.append(b)
sc_n = pb.make_scenarios(cfg, 40, seed=11, sigma=pb.SIGMA_NARROW)
sc_w = pb.make_scenarios(cfg, 21, seed=12, sigma=pb.SIGMA_WIDE)
x = np.concatenate([np.array(xk), sc_n.x_init, sc_w.x_init])
xs = np.concatenate([np.array(xsk), sc_n.ref_states, sc_w.ref_states])
us = np.concatenate([np.array(usk), sc_n.ref_inputs, sc_w.ref_inputs])
save("solves_N40_tracking.npz", cfg, x, xs, us, oracle.solve_batch(cfg, x, xs, us))

# 2. horizons 10 and 100 (config 3 end points), windows that run past the trajectory end included
for N in (10, 100):
    c = tracking_preset(N)
    sc = pb.make_scenarios(c, 16, seed=100 + N)
    save(f"solves_N{N}_tracking.npz", c, sc.x_init, sc.ref_states, sc.ref_inputs,
         oracle.solve_batch(c, sc.x_init, sc.ref_states, sc.ref_inputs))

# 3. NMPC preset (N=30, loose tolerances), cold and warm-started from the shifted cold solution
c = nmpc_preset(30)
sc = pb.make_scenarios(c, 16, seed=31, families=False)
r = oracle.solve_batch(c, sc.x_init, sc.ref_states, sc.ref_inputs)
zw = pb.shift_warm_start(r["z"], 30, reference_bug=True)
k2 = np.minimum(sc.k_index + 1, 400)
xs2, us2 = pb.windows_batch(S, U, k2, 30)
x2 = r["z"][:, 8:14].copy()  # nominal plant: next state = predicted x_1
r2 = oracle.solve_batch(c, x2, xs2, us2, z_warm=zw)
save("solves_N30_nmpc_cold.npz", c, sc.x_init, sc.ref_states, sc.ref_inputs, r)
save("solves_N30_nmpc_warm.npz", c, x2, xs2, us2, r2, z_warm=zw)
