"""Quick device-resident timing of the solve (development aid; bench.py is the contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset
from car_trailer_mpc_b200 import problem as pb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = tracking_preset(N); cfg.max_iter = 200
sc = pb.make_scenarios(cfg, B)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
x = torch.from_numpy(sc.x_init).to(dev); xs = torch.from_numpy(sc.ref_states).to(dev); us = torch.from_numpy(sc.ref_inputs).to(dev)
print("fp64 peak GFLOP/s:", s.measure_fp64_peak())
for _ in range(2):
    r = s.solve(x, xs, us)
torch.cuda.synchronize()
ts = []
for _ in range(steps):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = s.solve(x, xs, us); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
it = r["iters"].cpu().numpy(); st = r["status"].cpu().numpy()
print(f"B={B} N={N} ms/step={np.mean(ts):.3f} (min {np.min(ts):.3f}) solves/s={B/np.mean(ts)*1e3:.3e} iters mean={it.mean():.2f} max={it.max()} status={np.bincount(st, minlength=6)}")
