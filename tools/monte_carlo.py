"""Closed-loop Monte Carlo (size selectable): S scenarios x a full parking episode, sharded by
scenario id over the GPUs of one box (torchrun), metrics gathered / reduced with NCCL.

  python tools/monte_carlo.py --scenarios 65536 --t-sim 40 [--nominal] [--horizon 40]
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/monte_carlo.py --scenarios 1048576
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from car_trailer_mpc_b200 import BatchSolver, closed_loop as cl, problem as pb, sharding, tracking_preset

ap = argparse.ArgumentParser()
ap.add_argument("--scenarios", type=int, default=65536)
ap.add_argument("--horizon", type=int, default=40)
ap.add_argument("--t-sim", type=float, default=40.0)
ap.add_argument("--nominal", action="store_true", help="disturbances off")
ap.add_argument("--sigma0", type=float, default=0.02, help="std of the initial-state perturbation around S[0]")
ap.add_argument("--seed", type=int, default=2025)
ap.add_argument("--out", default="")
ap.add_argument("--device-loop", action="store_true", help="run whole episodes inside one persistent kernel (ttmpc_episode_batch)")
ap.add_argument("--max-iter", type=int, default=100)
args = ap.parse_args()

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

cfg = tracking_preset(args.horizon); cfg.max_iter = args.max_iter
S, U = pb.load_reference_trajectory(dt=cfg.dt)
lo, hi = sharding.shard_range(args.scenarios, rank, world)
ids = torch.arange(lo, hi, device=dev, dtype=torch.int64)
x0 = torch.as_tensor(S[0], device=dev)[None] + args.sigma0 * cl.counter_normal(args.seed, 10**6, ids, 6)
lb, ub = torch.tensor(list(cfg.x_lb), device=dev), torch.tensor(list(cfg.x_ub), device=dev)
x0[:, 2:] = torch.minimum(torch.maximum(x0[:, 2:], lb[2:] + 1e-3), ub[2:] - 1e-3)
solver = BatchSolver(cfg, local)
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.time()
dist_params = None if args.nominal else cl.DEFAULT_DISTURBANCE
if args.device_loop:
    ks = pb.time_indices(args.t_sim, cfg.dt)
    r = solver.episodes(x0.contiguous(), S, U, ks, dist_params, variant="tracking", seed=args.seed, scenario_ids=ids)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    wall = time.time() - t0
    rows = r["metrics"]
    out = {"steps": len(ks), "jackknife": rows[:, 4] > 0.5, "failures": rows[:, 5], "max_abs_psi": rows[:, 3],
           "max_iters": rows[:, 6]}
else:
    out = cl.simulate_batch(solver, S, U, x0.contiguous(), args.t_sim, cfg.dt, dist_params, seed=args.seed, scenario_ids=ids)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    wall = time.time() - t0
    rows = torch.stack([out["distance_error"], out["heading_error"].abs(), out["hitch_error"].abs(), out["max_abs_psi"],
                        out["jackknife"].double(), out["failures"].double(), out["mean_iters"], out["rms_tracking_error"]], 1)
# LQR score of the final state (LQR_cost.py:7-41, simulation.py:563): P from the DARE at the goal linearisation is the same
# for all scenarios -- solved once on the host, x'Px evaluated on the device, gathered as a ninth metric column
goal = S[-1]
try:
    P = torch.as_tensor(cl.lqr_riccati(cfg, cfg.Qm(), cfg.Rm(), goal, np.zeros(2)), device=dev)
    dxf = (r["final_state"] if args.device_loop else out["final_state"]) - torch.as_tensor(goal, device=dev)[None]
    lqr = ((dxf @ P) * dxf).sum(1)
except Exception as exc:  # the goal has v ~ -0.01: (A, B) is barely controllable and the DARE may fail
    if rank == 0:
        print(f"[monte_carlo] LQR score unavailable: {exc}", file=sys.stderr)
    lqr = torch.full((rows.shape[0],), float("nan"), dtype=torch.float64, device=dev)
rows = torch.cat([rows, lqr[:, None]], 1)
allrows = sharding.gather_rows(rows, args.scenarios)          # NCCL all-gather of the per-scenario metric rows
red = sharding.reduce_metrics({"jackknife": float(out["jackknife"].sum()), "failures": float(out["failures"].sum()),
                               "psi_max": float(out["max_abs_psi"].max()), "iters_max": float(out["max_iters"].max())})
if rank == 0:
    a = allrows.cpu().numpy()
    steps = out["steps"]
    q = lambda c: [float(np.percentile(a[:, c], p)) for p in (50, 99, 100)]
    summary = {"scenarios": args.scenarios, "gpus": world, "steps": steps, "horizon": args.horizon, "disturbances": not args.nominal, "device_loop": bool(args.device_loop),
               "solves": args.scenarios * steps, "wall_s": wall, "solves_per_s": args.scenarios * steps / wall,
               "jackknife_rate": red["jackknife"] / args.scenarios, "solver_failures": red["failures"], "max_abs_psi": red["psi_max"],
               "max_iters": red["iters_max"], "distance_error_p50_p99_max": q(0), "heading_error_p50_p99_max": q(1),
               "hitch_error_p50_p99_max": q(2), "mean_iters_p50_p99_max": q(6), "rms_tracking_error_p50_p99_max": q(7),
               "lqr_distance_p50_p99_max": q(8),
               "checksum_first_1024": float(a[:1024].sum())}
    print(json.dumps(summary))
    if args.out:
        json.dump(summary, open(args.out, "w"), indent=1)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
