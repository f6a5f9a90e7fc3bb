"""Opt-in geometric start of the OBCA duals (TTMPC_OBCA_GEOMETRIC_START) against the reference start on the config-4 batch:
times, iteration counts, convergence, and how often the two starts end in different local solutions.  usage: obca_geo_start_probe.py B"""
import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset, Obstacles, parking_lot_obstacles
from car_trailer_mpc_b200 import problem as pb
B = int(sys.argv[1]); N = 50
cfg = tracking_preset(N); cfg.max_iter = 300
S, U = pb.load_reference_trajectory()
rng = np.random.default_rng(20251018)
ks = rng.integers(0, 341, B).astype(np.int32)
lb = np.array(cfg.x_lb[:]); ub = np.array(cfg.x_ub[:])
x0 = S[ks] + rng.normal(0, 0.002, (B, 6)); x0[:, 2:] = np.clip(x0[:, 2:], lb[2:] + 1e-3, ub[2:] - 1e-3)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
args = (torch.from_numpy(x0).to(dev), torch.from_numpy(ks).to(dev), torch.from_numpy(S).to(dev), torch.from_numpy(U).to(dev))
out = {}
for geo in (False, True):
    obs = Obstacles.from_list(parking_lot_obstacles(), geometric_start=geo)
    s.solve_obca_shared(obs, *args); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = s.solve_obca_shared(obs, *args); e1.record(); torch.cuda.synchronize()
    out[geo] = {k: v.cpu().numpy() for k, v in r.items()}
    st = out[geo]["status"]
    print("geo" if geo else "ref", "ms %.1f" % e0.elapsed_time(e1), "converged %.4f" % (st <= 1).mean(), "iters mean %.1f max %d" % (out[geo]["iters"].mean(), out[geo]["iters"].max()), "hist", np.bincount(st, minlength=6))
a, b = out[False], out[True]
ok = (a["status"] == 0) & (b["status"] == 0)
print("both converged", ok.sum(), "max |du0| %.2e  max |dz| %.2e  max rel dJ %.2e" % (np.abs(a["u0"][ok] - b["u0"][ok]).max(), np.abs(a["z"][ok] - b["z"][ok]).max(), (np.abs(a["obj"][ok] - b["obj"][ok]) / np.abs(a["obj"][ok])).max()))
dz = np.abs(a["z"][ok] - b["z"][ok]).max(axis=1)
print("problems with |dz| > 1e-6:", int((dz > 1e-6).sum()), " > 1e-4:", int((dz > 1e-4).sum()))
print("ref-only converged", int(((a["status"] <= 1) & (b["status"] > 1)).sum()), "geo-only converged", int(((a["status"] > 1) & (b["status"] <= 1)).sum()))
if (dz > 1e-6).any():
    idx = np.where(ok)[0][dz > 1e-6]
    ja, jb = a["obj"][idx], b["obj"][idx]
    passage = (ks[idx] + N >= 117 - 35) & (ks[idx] <= 117)
    print("differing solutions: geo objective lower in %d, higher in %d of %d; median J ref %.4g geo %.4g; windows reaching the passage: %d"
          % (int((jb < ja * (1 - 1e-9)).sum()), int((jb > ja * (1 + 1e-9)).sum()), len(idx), np.median(ja), np.median(jb), int(passage.sum())))
    print("constraint violation of the differing ones: ref max %.2e geo max %.2e" % (a["kkt"][idx, 1].max(), b["kkt"][idx, 1].max()))
