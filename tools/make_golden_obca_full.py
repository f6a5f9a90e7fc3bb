"""Generates tests/golden/obca_cases_full.npz: OBCA problems at the REFERENCE'S OWN SIZE (simulation.py:390: horizon 50, and
horizon 40; all 11 rectangles of obstacles.json, one case with a twelfth obstacle that blocks the path so that
collision rows are active, one in which the solve recovers from an exhausted line search) + the oracle's solutions with its block-tridiagonal linear solver
(oracle/obca_oracle.py, linear_solver="banded": same algorithm as the dense path -- tests/test_obca_cpu.py checks that
both walk the same iterates -- but it finishes at this size).  The reference itself cannot be run here (no CasADi/Ipopt).
Re-run with:  python tools/make_golden_obca_full.py   (about ten minutes)"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from car_trailer_mpc_b200 import tracking_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402
from car_trailer_mpc_b200.config import parking_lot_obstacles  # noqa: E402
from oracle import obca_oracle as ob  # noqa: E402

S, U = pb.load_reference_trajectory()
W1, W2 = 3.05, 2.95
lot = parking_lot_obstacles()
SEED = 41


def blocking_obstacle(k0, stage, gap):
    """A 1 m square next to the vehicle body at `stage` of the window, `gap` metres from its side (< d_min)."""
    x, y, th = S[k0 + stage, :3]
    pc = np.array([x + np.cos(th) * 7.05 / 2, y + np.sin(th) * 7.05 / 2])
    n = np.array([-np.sin(th), np.cos(th)])
    half_diag = 0.5 * (abs(np.cos(th)) + abs(np.sin(th)))
    c = pc + n * (W1 / 2 + gap + half_diag)
    return dict(center=(float(c[0]), float(c[1])), width=1.0, height=1.0)


# a start for which the line search runs out of backtracking steps on the way (problem 26 of tools/obca_bench.py's
# scenario stream): the recovery of DESIGN.md section 3b is taken, by the oracle and by the kernel
DX_RECOVERY = np.array([-0.00151298, -0.00038422, 0.00160258, -0.00333941, 0.00037641, -0.00234964])
cases = [  # (name, N, k0, obstacles, sigma or explicit offset of x_init from the reference)
    ("n50_k200_11obs", 50, 200, lot, 0.002),
    ("n40_k250_11obs", 40, 250, lot, 0.002),
    ("n50_k60_12obs_blocked", 50, 60, lot + [blocking_obstacle(60, 10, 0.15)], 0.002),
    ("n50_k144_11obs_recovery", 50, 144, lot, DX_RECOVERY),
]
if __name__ == "__main__":
    only = sys.argv[1:]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    path = os.path.join(ROOT, "tests", "golden", "obca_cases_full.npz")
    out = dict(np.load(path)) if os.path.exists(path) and only else {}
    rng = np.random.default_rng(SEED)
    for name, N, k0, obst, sig in cases:
        cfg = tracking_preset(N)
        rs, ru = pb.window(S, U, k0, N)
        noise = rng.normal(0, 0.002, 6)  # always drawn: the cases' starts do not depend on which ones are regenerated
        x0 = rs[0] + (noise if np.isscalar(sig) else sig)
        if only and name not in only:
            continue
        nlp = ob.ObcaNlp(N, cfg.dt, cfg.L1, cfg.L2, cfg.M, W1, W2, cfg.Qm(), np.array(cfg.R[:]).reshape(2, 2), list(cfg.x_lb),
                         list(cfg.x_ub), list(cfg.u_lb), list(cfg.u_ub), obst)
        t = time.time()
        r = ob.solve(nlp, x0, rs, ru, tol=cfg.tol, acc_tol=cfg.acceptable_tol, acc_iter=cfg.acceptable_iter, max_iter=400,
                     linear_solver="banded")
        act = sum(int(abs(r["w"][nlp.isl(k, j)][0]) < 1e-6) for k in range(N + 1) for j in range(nlp.P))
        print(f"{name}: status {r['status']} iters {r['iters']} obj {r['obj']:.9e} active rows {act} ({time.time() - t:.1f} s)", flush=True)
        out[name + "/horizon"] = N
        out[name + "/x_init"] = x0
        out[name + "/ref_states"] = rs
        out[name + "/ref_inputs"] = ru
        out[name + "/rects"] = np.array([[o["center"][0], o["center"][1], o["width"], o["height"]] for o in obst])
        out[name + "/states"] = r["states"]
        out[name + "/inputs"] = r["inputs"]
        out[name + "/obj"] = r["obj"]
        out[name + "/iters"] = r["iters"]
        out[name + "/status"] = r["status"]
        out[name + "/active_rows"] = act
        out[name + "/restarts"] = r["restarts"]
    out["generator"] = "tools/make_golden_obca_full.py"
    out["seed"] = SEED
    out["git_commit"] = commit
    np.savez_compressed(path, **out)
