"""Development aid: device-resident timing of the plain solve for each kernel flavour (lane kernel, team kernel with
8/16/32 lanes per problem) over a list of (B, N) cases, with statuses / iteration counts cross-checked between them.
Usage: python tools/team_probe.py "65536x40,4096x40,1x40" [steps] [lane,8,16,32,auto]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from car_trailer_mpc_b200 import BatchSolver, tracking_preset
from car_trailer_mpc_b200 import problem as pb

cases = [tuple(int(v) for v in c.split("x")) for c in (sys.argv[1] if len(sys.argv) > 1 else "65536x40,4096x40,1x40").split(",")]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["lane", "8", "16", "32", "auto"]
dev = torch.device("cuda:0")
for B, N in cases:
    cfg = tracking_preset(N); cfg.max_iter = 200
    sc = pb.make_scenarios(cfg, B)
    x = torch.from_numpy(sc.x_init).to(dev); xs = torch.from_numpy(sc.ref_states).to(dev); us = torch.from_numpy(sc.ref_inputs).to(dev)
    base = None
    for mode in modes:
        os.environ.pop("TTMPC_KERNEL", None); os.environ.pop("TTMPC_TEAM_LANES", None)
        os.environ.pop("TTMPC_TEAM_ORDER", None)
        if mode == "lane":
            os.environ["TTMPC_KERNEL"] = "lane"
        elif mode.endswith("o"):   # e.g. "16o": team kernel, hardest-first order
            os.environ["TTMPC_TEAM_LANES"] = mode[:-1]
            os.environ["TTMPC_TEAM_ORDER"] = "1"
        elif mode != "auto":
            os.environ["TTMPC_TEAM_LANES"] = mode
        s = BatchSolver(cfg, 0)
        for _ in range(2):
            r = s.solve(x, xs, us)
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = s.solve(x, xs, us); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        it = r["iters"].cpu().numpy(); st = r["status"].cpu().numpy(); u0 = r["u0"].cpu().numpy()
        if base is None:
            base = (it, st, u0)
        same = f"status_eq={np.array_equal(st, base[1])} iters_eq={(it == base[0]).mean():.4f} du0={np.abs(u0 - base[2]).max():.1e}"
        print(f"B={B} N={N} mode={mode:>4} lanes={s.last_solve_lanes()} ms={np.mean(ts):.3f} (min {np.min(ts):.3f}) "
              f"solves/s={B / np.mean(ts) * 1e3:.3e} iters={it.mean():.2f}/{it.max()} {same}", flush=True)
        s.close()
