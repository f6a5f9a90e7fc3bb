"""How often the plain solve ends without converging, and how: the headline batch (narrow sigma), the wide-sigma batch of
SURVEY 8(d) and 10x that, 65 536 problems each, GPU (lane kernel) -- and the CPU oracle on the first 2 048 of each for the
same counts (the oracle implements the same rules; the reference's Ipopt, which would enter feasibility restoration where
these line searches give up, cannot run in this image).  usage: wide_sigma_status.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset
from car_trailer_mpc_b200 import problem as pb
from oracle import oracle

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = tracking_preset(40); cfg.max_iter = 200
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
names = ["converged", "acceptable", "max_iter", "line search", "numeric", "x0 outside bounds"]
for label, sig in (("narrow", pb.SIGMA_NARROW), ("wide", pb.SIGMA_WIDE), ("10 x wide", 10.0 * np.asarray(pb.SIGMA_WIDE))):
    sc = pb.make_scenarios(cfg, B, seed=77, sigma=sig)
    r = s.solve(torch.from_numpy(sc.x_init).to(dev), torch.from_numpy(sc.ref_states).to(dev), torch.from_numpy(sc.ref_inputs).to(dev), want_z=False)
    st = r["status"].cpu().numpy(); it = r["iters"].cpu().numpy()
    m = min(B, 2048)
    o = oracle.solve_batch(cfg, sc.x_init[:m], sc.ref_states[:m], sc.ref_inputs[:m])
    print(f"{label:10s} GPU {dict(zip(names, np.bincount(st, minlength=6).tolist()))} iters mean {it.mean():.2f} max {it.max()} | "
          f"oracle on the first {m}: {np.bincount(o['status'], minlength=6).tolist()}, GPU on the same: {np.bincount(st[:m], minlength=6).tolist()}")
