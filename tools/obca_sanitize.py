"""One small obstacle-aware solve per kernel flavour, for compute-sanitizer runs (memcheck / racecheck / synccheck):
usage: compute-sanitizer --tool racecheck python tools/obca_sanitize.py [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset, Obstacles, parking_lot_obstacles
from car_trailer_mpc_b200 import problem as pb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 12
B = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = tracking_preset(N); cfg.max_iter = 60
S, U = pb.load_reference_trajectory()
ks = np.array([100, 150, 200, 50, 250][:B], dtype=np.int32)
x0 = S[ks].copy()
obs = Obstacles.from_list(parking_lot_obstacles())
dev = torch.device("cuda:0")
for env, name in (({"TTMPC_OBCA_CLUSTER": "4"}, "cluster of 4"), ({"TTMPC_OBCA_CLUSTER": "0"}, "CTA per problem"),
                  ({"TTMPC_OBCA_WIDE_MAX": "0"}, "warp per problem")):
    for k in ("TTMPC_OBCA_CLUSTER", "TTMPC_OBCA_WIDE_MAX"):
        os.environ.pop(k, None)
    os.environ.update(env)
    s = BatchSolver(cfg, 0)
    r = s.solve_obca_shared(obs, torch.from_numpy(x0).to(dev), torch.from_numpy(ks).to(dev), torch.from_numpy(S).to(dev),
                            torch.from_numpy(U).to(dev))
    torch.cuda.synchronize()
    print(name, "status", r["status"].cpu().numpy(), "iters", r["iters"].cpu().numpy(), {k: v for k, v in s.kernel_launches().items() if v})
    s.close()
