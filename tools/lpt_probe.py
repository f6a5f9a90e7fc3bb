"""Upper bound of longest-processing-time-first scheduling: re-submit the batch sorted by the true iteration count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset, problem as pb
B = 65536
cfg = tracking_preset(40); cfg.max_iter = 200
sc = pb.make_scenarios(cfg, B)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
def run(order, label):
    x = torch.from_numpy(sc.x_init[order]).to(dev); xs = torch.from_numpy(sc.ref_states[order]).to(dev); us = torch.from_numpy(sc.ref_inputs[order]).to(dev)
    for _ in range(2): r = s.solve(x, xs, us)
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = s.solve(x, xs, us); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{label:28s} {np.mean(ts):7.3f} ms")
    return r["iters"].cpu().numpy()
it = run(np.arange(B), "arrival order")
run(np.argsort(-it, kind="stable"), "LPT (true iterations)")
run(np.argsort(it, kind="stable"), "SPT (worst case)")
rng = np.random.default_rng(0); run(rng.permutation(B), "random permutation")
# uniform work: every problem identical (5 iterations) -> pure quantisation of B / lanes
one = np.zeros(B, dtype=np.int64) + int(np.argmin(it))
run(one, "identical problems")
