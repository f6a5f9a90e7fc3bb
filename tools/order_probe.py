"""Scheduling-order experiments for the solve kernel (development aid): run with TTMPC_NO_ORDER=1 so that the library
keeps the submitted order, and submit the benchmark batch sorted by different hardness scores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset, problem as pb
B = 65536
cfg = tracking_preset(40); cfg.max_iter = 200
sc = pb.make_scenarios(cfg, B, seed=20251018)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
def run(order, label):
    x = torch.from_numpy(sc.x_init[order]).to(dev); xs = torch.from_numpy(sc.ref_states[order]).to(dev); us = torch.from_numpy(sc.ref_inputs[order]).to(dev)
    for _ in range(2): r = s.solve(x, xs, us)
    torch.cuda.synchronize(); ts = []
    for _ in range(6):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(x, xs, us); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{label:44s} {np.mean(ts):7.3f} ms", flush=True)
    return r["iters"].cpu().numpy()
it = run(np.arange(B), "arrival order")
lo = np.array(list(cfg.x_lb) + list(cfg.u_lb)); up = np.array(list(cfg.x_ub) + list(cfg.u_ub))
two = np.isfinite(lo) & np.isfinite(up)
m = 0.05 * np.where(two, up - lo, 1.0)
xs, us = sc.ref_states, sc.ref_inputs
nx = ((((xs[:, 1:, :] - lo[:6]) < m[:6]) | ((up[:6] - xs[:, 1:, :]) < m[:6])) & two[:6]).sum((1, 2))
nu = ((((us - lo[6:]) < m[6:]) | ((up[6:] - us) < m[6:])) & two[6:]).sum((1, 2))
X = xs.copy(); X[:, 0, :] = sc.x_init
th0 = np.abs(pb.dynamics_defect(cfg, X, us)).sum((1, 2))
near = nx + nu
c4 = np.where(near >= 24, 3, np.where(near >= 8, 2, np.where(near >= 1, 1, 0)))
c4b = c4.copy(); c4b[(th0 > 0.5) & (c4 < 2)] = 2
s8 = np.zeros(B, int)
s8[nx >= 1] = 2; s8[nx >= 4] = 4; s8[nx >= 7] = 6; s8[nx >= 10] = 7; s8[nx >= 20] = 5; s8[(nx == 0) & (nu >= 1)] = 1; s8[(th0 > 0.5) & (s8 < 3)] = 3
rng = np.random.default_rng(0)
def by(score):  # descending score, random within a class (as the atomics of the order kernel do)
    return np.lexsort((rng.random(B), -score))
run(by(c4), "4 classes (old)")
run(by(c4b), "4 classes + theta0 (shipped)")
run(by(s8), "8 classes")
run(by(it), "true iterations, random within")
run(np.argsort(-it, kind="stable"), "true iterations, stable")
# interleave: hardest problems spread over CTAs instead of packed into the first ones
o = by(it); L = 37888
run(np.concatenate([o[:L].reshape(-1, 148).T.ravel(), o[L:]]), "true iterations, first wave transposed")
