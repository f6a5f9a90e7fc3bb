"""Generates tests/golden/planner_cases.npz: problems of the offline planner's NLP (trajectory_optimization.py: goal cost
with terminal weight 100 Q, final-state box +-1e-2, all collision rows; settings of trajectory_animation.py:41-79) + the
oracle's solutions (oracle/obca_oracle.py, block-tridiagonal linear solver).  Start, goal and the initial trajectory are
cut from the reference's stored planner output data/state_traj.txt (dt = 0.1): the stored poses with steering angle, speed
and inputs at zero -- what `_hybrid_a_star_initial_trajectory` (:227-274) builds from a waypoint path.  The last case is
the reference's own size: horizon 200, the 11 rectangles of obstacles.json (37 k variables).  The reference itself cannot
be run here (no CasADi/Ipopt).   python tools/make_golden_planner.py [case ...]   (the full-size case takes ~15 minutes)"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from car_trailer_mpc_b200 import planner_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402
from car_trailer_mpc_b200.config import parking_lot_obstacles  # noqa: E402
from oracle import obca_oracle as ob  # noqa: E402

SRAW = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt")).T  # [201, 6], the planner's stored output
LOT = parking_lot_obstacles()
W1, W2 = 3.05, 2.95
TERM_W, TERM_BOX = 100.0, 1e-2


def case_data(N, k0, nob):
    x0 = SRAW[k0].copy()
    goal = SRAW[k0 + N].copy()
    goal[4:] = 0.0  # trajectory_animation.py:92-93: steering angle and speed of the goal are zero
    obst = sorted(LOT, key=lambda o: abs(o["center"][0] - SRAW[k0 + N // 2, 0]))[:nob]
    guess = np.zeros((N + 1, 6))
    guess[:, :4] = SRAW[k0:k0 + N + 1, :4]
    return x0, goal, obst, guess


cases = [("plan_n10_k100_2obs", 10, 100, 2), ("plan_n40_k150_11obs", 40, 150, 11), ("plan_n200_11obs", 200, 0, 11)]
if __name__ == "__main__":
    only = sys.argv[1:]
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    path = os.path.join(ROOT, "tests", "golden", "planner_cases.npz")
    out = dict(np.load(path)) if os.path.exists(path) else {}
    for name, N, k0, nob in cases:
        if only and name not in only:
            continue
        cfg = planner_preset(N)
        x0, goal, obst, guess = case_data(N, k0, nob)
        nlp = ob.ObcaNlp(N, cfg.dt, cfg.L1, cfg.L2, cfg.M, W1, W2, cfg.Qm(), np.array(cfg.R[:]).reshape(2, 2), list(cfg.x_lb),
                         list(cfg.x_ub), list(cfg.u_lb), list(cfg.u_ub), obst, terminal_weight=TERM_W, terminal_box=(goal, TERM_BOX))
        t = time.time()
        r = ob.solve(nlp, x0, np.tile(goal, (N + 1, 1)), np.zeros((N, 2)), tol=cfg.tol, acc_tol=cfg.acceptable_tol,
                     acc_iter=cfg.acceptable_iter, max_iter=1000, linear_solver="banded", guess=(guess, np.zeros((N, 2))))
        print(f"{name}: status {r['status']} iters {r['iters']} obj {r['obj']:.9e} restarts {r['restarts']} ({time.time() - t:.1f} s)", flush=True)
        out[name + "/horizon"] = N
        out[name + "/x_init"] = x0
        out[name + "/goal"] = goal
        out[name + "/guess"] = guess
        out[name + "/rects"] = np.array([[o["center"][0], o["center"][1], o["width"], o["height"]] for o in obst])
        out[name + "/states"] = r["states"]
        out[name + "/inputs"] = r["inputs"]
        out[name + "/obj"] = r["obj"]
        out[name + "/iters"] = r["iters"]
        out[name + "/status"] = r["status"]
        out["generator"] = "tools/make_golden_planner.py"
        out["git_commit"] = commit
        np.savez_compressed(path, **out)
