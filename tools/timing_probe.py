import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset, problem as pb, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfg = tracking_preset(40); cfg.max_iter = 200
sc = pb.make_scenarios(cfg, B)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
L = _lib.load()
x = torch.from_numpy(sc.x_init).to(dev); xs = torch.from_numpy(sc.ref_states).to(dev); us = torch.from_numpy(sc.ref_inputs).to(dev)
for _ in range(2): s.solve(x, xs, us)
torch.cuda.synchronize()
L.ttmpc_debug_timing(None, 1)
r = s.solve(x, xs, us); torch.cuda.synchronize()
out = (ctypes.c_ulonglong * 8)()
L.ttmpc_debug_timing(out, 0)
names = ["loop tail (unpack, result scalars)", "ticket fetch", "-", "barrier 1", "backward half", "barrier 2", "step half (fwd+trials)", "active lane-rounds"]
tot = sum(out[i] for i in range(7))
for n, v in zip(names, out): print(f"{n:36s} {v:>16d}  {100*v/tot:5.1f}%")
print("warp-rounds:", "active lanes per warp-round = n/a")
