"""Batch driver: sweeps solver configurations over synthetic scenario batches and logs one CSV row per run.

New code (the reference has no MPC batch driver); its command-line/CSV shape follows the
reference's sweep tool ``python-files/compare_sweep.py`` (argparse lists -> cartesian product -> header + one row per
combination, ``--output`` / ``--append``, :59-152) and its scenario families are the poses of ``test_cases.json``.

  python -m car_trailer_mpc_b200.batch_driver --horizons 30 40 50 --batches 4096 --sigmas narrow wide \
      --presets tracking nmpc --output sweep_mpc.csv
"""
from __future__ import annotations

import argparse
import csv
import sys
import time
from itertools import product
from pathlib import Path

import numpy as np

from . import problem as pb
from .config import STATUS_NAMES, nmpc_preset, tracking_preset

HEADER = ["preset", "horizon", "batch", "sigma", "families", "device_ms", "solves_per_s", "mean_iters", "max_iters",
          "frac_converged", "frac_acceptable", "frac_failed", "kkt_dual_p99", "kkt_viol_max", "kkt_compl_p99",
          "u0_a_mean", "u0_w_mean"] + [f"n_family_{i}" for i in range(-1, 7)]


def run_one(preset: str, horizon: int, batch: int, sigma: str, families: bool, device: int, repeats: int, max_iter: int):
    import torch

    from .solver import BatchSolver

    cfg = tracking_preset(horizon) if preset == "tracking" else nmpc_preset(horizon)
    cfg.max_iter = max_iter
    sc = pb.make_scenarios(cfg, batch, sigma=pb.SIGMA_NARROW if sigma == "narrow" else pb.SIGMA_WIDE, families=families)
    dev = torch.device("cuda", device)
    s = BatchSolver(cfg, device)
    x, xs, us = (torch.from_numpy(a).to(dev) for a in (sc.x_init, sc.ref_states, sc.ref_inputs))
    with torch.cuda.device(dev):  # events and the current stream belong to --device, not to the process's default device
        r = s.solve(x, xs, us, want_z=False)  # warm-up
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(repeats):
            r = s.solve(x, xs, us, want_z=False)
        e1.record()
        torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / repeats
    st = r["status"].cpu().numpy()
    it = r["iters"].cpu().numpy()
    kkt = r["kkt"].cpu().numpy()
    u0 = r["u0"].cpu().numpy()
    fam = np.bincount(sc.family + 1, minlength=8)
    return [preset, horizon, batch, sigma, int(families), f"{ms:.3f}", f"{batch / ms * 1e3:.1f}", f"{it.mean():.3f}", int(it.max()),
            f"{(st == 0).mean():.6f}", f"{(st == 1).mean():.6f}", f"{(st > 1).mean():.6f}",
            f"{np.percentile(kkt[:, 0], 99):.3e}", f"{kkt[:, 1].max():.3e}", f"{np.percentile(kkt[:, 2], 99):.3e}",
            f"{u0[:, 0].mean():.6f}", f"{u0[:, 1].mean():.6f}"] + [int(v) for v in fam]


def main(argv=None):
    ap = argparse.ArgumentParser(description="Sweep the batched NMPC solver and log throughput / convergence per configuration.")
    ap.add_argument("--presets", nargs="+", default=["tracking"], choices=["tracking", "nmpc"])
    ap.add_argument("--horizons", nargs="+", type=int, default=[40])
    ap.add_argument("--batches", nargs="+", type=int, default=[4096])
    ap.add_argument("--sigmas", nargs="+", default=["narrow"], choices=["narrow", "wide"])
    ap.add_argument("--no-families", action="store_true", help="do not move the trajectory to the test_cases.json start poses")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--max-iter", type=int, default=200)
    ap.add_argument("--output", type=Path, default=Path("sweep_mpc.csv"))
    ap.add_argument("--append", action="store_true", help="Append to the CSV instead of overwriting.")
    args = ap.parse_args(argv)

    mode = "a" if args.append and args.output.exists() else "w"
    args.output.parent.mkdir(parents=True, exist_ok=True)
    with args.output.open(mode, newline="", encoding="utf-8") as f:
        w = csv.writer(f)
        if mode == "w":
            w.writerow(HEADER)
        for preset, horizon, batch, sigma in product(args.presets, args.horizons, args.batches, args.sigmas):
            t0 = time.time()
            row = run_one(preset, horizon, batch, sigma, not args.no_families, args.device, args.repeats, args.max_iter)
            w.writerow(row)
            f.flush()
            print(f"Logged preset={preset}, N={horizon}, B={batch}, sigma={sigma} -> {row[6]} solves/s, "
                  f"iters {row[7]}/{row[8]}, converged {row[9]} ({time.time() - t0:.1f}s)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
