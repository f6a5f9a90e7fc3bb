"""Generates tests/golden/obca_cases.npz: seeded OBCA problems (config 4 of SURVEY 8(d), small sizes) + the DENSE oracle's
solutions (oracle/obca_oracle.py; the reference itself cannot be run here -- CasADi/Ipopt are not installable).
Re-run with:  python tools/make_golden_obca.py   (about 3 minutes)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from car_trailer_mpc_b200 import tracking_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402
from car_trailer_mpc_b200.config import parking_lot_obstacles  # noqa: E402
from oracle import obca_oracle as ob  # noqa: E402

S, U = pb.load_reference_trajectory()
W1, W2 = 3.05, 2.95
lot = parking_lot_obstacles()


def nearest(k0, n):
    return sorted(lot, key=lambda o: abs(o["center"][0] - S[k0, 0]))[:n]


def blocking_obstacle(k0, N, stage, gap):
    """A 1 m square next to the vehicle body at `stage` of the window, `gap` metres from its side: closer than d_min, so
    the tracking solution has to give way (active collision rows)."""
    x, y, th = S[k0 + stage, :3]
    pc = np.array([x + np.cos(th) * 7.05 / 2, y + np.sin(th) * 7.05 / 2])
    n = np.array([-np.sin(th), np.cos(th)])
    half_diag = 0.5 * (abs(np.cos(th)) + abs(np.sin(th)))  # support of the axis-aligned unit square along n
    c = pc + n * (W1 / 2 + gap + half_diag)
    return dict(center=(float(c[0]), float(c[1])), width=1.0, height=1.0)


cases = [  # (name, N, k0, obstacles, sigma)
    ("n6_k300_2obs", 6, 300, nearest(300, 2), 0.002),
    ("n3_k340_11obs", 3, 340, nearest(340, 11), 0.002),
    ("n8_k200_11obs", 8, 200, nearest(200, 11), 0.002),
    ("n30_k290_2obs", 30, 290, nearest(290, 2), 0.002),
    ("n12_k60_blocked", 12, 60, [blocking_obstacle(60, 12, 6, 0.15)], 0.002),
    ("n12_k200_blocked3", 12, 200, [blocking_obstacle(200, 12, 6, 0.15)] + nearest(200, 2), 0.002),
]
out = {}
rng = np.random.default_rng(4)
for name, N, k0, obst, sig in cases:
    cfg = tracking_preset(N)
    rs, ru = pb.window(S, U, k0, N)
    x0 = rs[0] + rng.normal(0, sig, 6)
    nlp = ob.ObcaNlp(N, cfg.dt, cfg.L1, cfg.L2, cfg.M, W1, W2, cfg.Qm(), np.array(cfg.R[:]).reshape(2, 2), list(cfg.x_lb),
                     list(cfg.x_ub), list(cfg.u_lb), list(cfg.u_ub), obst)
    t = time.time()
    r = ob.solve(nlp, x0, rs, ru, tol=cfg.tol, acc_tol=cfg.acceptable_tol, acc_iter=cfg.acceptable_iter, max_iter=cfg.max_iter)
    # active collision rows: slack of d0 within 1e-6 of its bound
    act = sum(int(abs(r["w"][nlp.isl(k, j)][0]) < 1e-6) for k in range(N + 1) for j in range(nlp.P))
    print(f"{name}: status {r['status']} iters {r['iters']} obj {r['obj']:.9e} active rows {act} ({time.time() - t:.1f} s)")
    assert r["status"] == 0
    out[name + "/horizon"] = N
    out[name + "/x_init"] = x0
    out[name + "/ref_states"] = rs
    out[name + "/ref_inputs"] = ru
    out[name + "/rects"] = np.array([[o["center"][0], o["center"][1], o["width"], o["height"]] for o in obst])
    out[name + "/states"] = r["states"]
    out[name + "/inputs"] = r["inputs"]
    out[name + "/obj"] = r["obj"]
    out[name + "/iters"] = r["iters"]
    out[name + "/active_rows"] = act
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "obca_cases.npz"), **out)
