import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset
from car_trailer_mpc_b200 import problem as pb
B, N = 65536, 40
cfg = tracking_preset(N); cfg.max_iter = 200
sc = pb.make_scenarios(cfg, B)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
x = torch.from_numpy(sc.x_init).to(dev); xs = torch.from_numpy(sc.ref_states).to(dev); us = torch.from_numpy(sc.ref_inputs).to(dev)
base = None
for i, wz in enumerate([False, True, True, False, True]):
    r = s.solve(x, xs, us, want_z=wz)
    it = r["iters"].cpu().numpy(); u0 = r["u0"].cpu().numpy(); st = r["status"].cpu().numpy()
    if base is None: base = (it, u0)
    d = np.nonzero(it != base[0])[0]
    print(f"run {i} want_z={wz} iters {it.mean():.4f}/{it.max()} status {np.bincount(st, minlength=6).tolist()} iters differing from run 0: {d.size} u0 diff {np.abs(u0-base[1]).max():.2e} first {d[:8].tolist()} {it[d[:8]].tolist()} vs {base[0][d[:8]].tolist()}")
