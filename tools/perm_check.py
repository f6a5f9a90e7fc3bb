"""Development aid: permutation invariance of the plain solve on the GPU (bit-identical u0 / iterations for any slot
assignment) + device-resident timing, for the library named by TTMPC_LIB.  usage: perm_check.py [B] [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hashlib
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset
from car_trailer_mpc_b200 import problem as pb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
os.environ["TTMPC_KERNEL"] = "lane"
cfg = tracking_preset(N); cfg.max_iter = 200
sc = pb.make_scenarios(cfg, B)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
x = torch.from_numpy(sc.x_init).to(dev); xs = torch.from_numpy(sc.ref_states).to(dev); us = torch.from_numpy(sc.ref_inputs).to(dev)
r = s.solve(x, xs, us, want_z=False)
u0 = r["u0"].cpu().numpy(); it = r["iters"].cpu().numpy()
bad = 0
for seed in range(3):
    perm = np.random.default_rng(seed).permutation(B); pt = torch.from_numpy(perm).to(dev)
    rp = s.solve(x[pt].contiguous(), xs[pt].contiguous(), us[pt].contiguous(), want_z=False)
    bad += int((rp["u0"].cpu().numpy() != u0[perm]).any(1).sum()) + int((rp["iters"].cpu().numpy() != it[perm]).sum())
ts = []
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); s.solve(x, xs, us, want_z=False); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"lib={os.environ.get('TTMPC_LIB','default')} B={B} N={N} mismatches={bad} ms={np.mean(ts):.3f} iters={it.mean():.3f}/{it.max()} "
      f"u0_sha={hashlib.sha256(u0.tobytes()).hexdigest()[:16]} lanes={s.last_solve_lanes()}")
