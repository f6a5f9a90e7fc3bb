"""Timing of the offline planner's NLP (ttmpc_plan_batch) at the reference's own size: horizon 200, dt 0.1, the 11
rectangles of obstacles.json, start / goal / initial poses from the stored planner output (development aid + profiles/).
usage: planner_bench.py [B] [--json out.json] [--cpu]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, planner_preset, Obstacles, parking_lot_obstacles
from car_trailer_mpc_b200 import problem as pb
ap = argparse.ArgumentParser(); ap.add_argument("B", type=int, nargs="?", default=1); ap.add_argument("--json", default=None)
ap.add_argument("--geo-start", action="store_true", help="TTMPC_OBCA_GEOMETRIC_START (opt-in, not the reference's starting duals)")
ap.add_argument("--cpu", action="store_true", help="also time the host build of the same core on one problem (one thread)")
a = ap.parse_args()
N = 200
S = np.loadtxt(os.path.join(pb.DATA_DIR, "state_traj.txt")).T
cfg = planner_preset(N); cfg.max_iter = 1000
x0, goal = S[0].copy(), S[N].copy(); goal[4:] = 0.0
zg = np.zeros(8 * N + 6)
for k in range(N + 1): zg[8 * k:8 * k + 4] = S[k, :4]
rng = np.random.default_rng(0)
X0 = x0[None] + np.concatenate([np.zeros((1, 6)), rng.normal(0, 1e-3, (a.B - 1, 6))])
obs = Obstacles.from_list(parking_lot_obstacles(), geometric_start=a.geo_start)
dev = torch.device("cuda:0")
s = BatchSolver(cfg, 0)
tx, tz = torch.from_numpy(X0).to(dev), torch.from_numpy(np.tile(zg, (a.B, 1))).to(dev)
ts = []
for i in range(4):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = s.plan(obs, tx, goal, 100.0, 1e-2, tz); e1.record(); torch.cuda.synchronize()
    if i: ts.append(e0.elapsed_time(e1))
st = r["status"].cpu().numpy(); it = r["iters"].cpu().numpy()
res = dict(B=a.B, N=N, obstacles=11, variables=8 * N + (N + 1) * 22 * 8, slacks=(N + 1) * 22 * 4, ms=float(np.mean(ts)), status_hist=np.bincount(st, minlength=6).tolist(),
           iters=it.tolist()[:8], obj=float(r["obj"][0]), kernels=s.kernel_launches())
if a.cpu:
    import emu
    t = time.time(); e = emu.obca_plan_batch(cfg, obs, X0[:1], goal, 100.0, 1e-2, zg[None], wide_warps=8); dt = time.time() - t
    res["cpu_port"] = dict(seconds=dt, cores=1, iters=int(e["iters"][0]), sample="the same problem, host build of the kernel's solver core (g++ -O2, one thread)",
                           max_dz_vs_gpu=float(np.abs(e["z"][0] - r["z"][0].cpu().numpy()).max()))
print(json.dumps(res))
if a.json: json.dump(res, open(a.json, "w"), indent=1)
