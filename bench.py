#!/usr/bin/env python
"""bench.py -- headline benchmark: batched NMPC solves/sec (horizon 40, 65 536 scenarios per step and GPU).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--horizon H]

N > 1 is launched by torchrun (one rank per GPU, NCCL); every rank owns an independent shard of B scenarios
(weak scaling, no data-path collective: the solve has no exchange step) and NCCL only gathers the per-scenario
first controls / status after each step, as BASELINE.json's north star prescribes.  Rank 0 prints ONE JSON line.

A "step" = one pass of the hot path over one batch of synthetic scenarios (SURVEY.md section 8(d) config 2 at
B = 65 536): perturbed initial states around the shipped reference trajectory, per-problem reference windows,
`MPCTrackingControl` preset (Ipopt defaults, tol 1e-8), cold start.

  value     whole-job solves/s with inputs already resident in HBM (CUDA events, max over ranks).
  e2e       same metric through the library's own host-pointer call (ttmpc_solve_batch, TTMPC_FLAG_HOST_POINTERS |
            TTMPC_FLAG_ASYNC_HOST: the library's three-stream copy-in | solve | copy-out pipeline) on page-locked HOST
            buffers: every step moves its x_init / reference windows in and the full decision vectors + u0 / status
            out inside the timed region; timed on the device (CUDA events on the library's copy streams).
            `e2e_compact` is the same through ttmpc_solve_batch_multi (x_init + window start + trajectory index in,
            u0 / iterations / status out: 80 B per problem instead of 5.3 KB).
  strong    the SAME 65 536 scenarios split over the ranks (strong scaling, automatic kernel choice) and
            `u0_checksum_global`, a checksum of all their first controls computed with one fixed kernel flavour, which
            must not depend on the number of GPUs.
  secondary configs 2, 3, 4 of BASELINE.json at reduced step counts (B = 4096 / N = 40, B = 1 latency, N = 100, OBCA).
  roofline  dominant kernel = ttmpc_solve_kernel.  It is bound by the non-tensor FP64 pipe / dependent-issue
            latency, not by HBM or tensor cores; `achieved` = algorithmic FP64 flop
            (mean_iters * (2500 N + 500) per solve) / kernel time, `peak` = the FP64 FMA peak measured in this run
            by the library's DFMA microbenchmark (MEASURED_PEAKS.json has no FP64 figure).  The HBM view
            (algorithmic bytes 5304 B/solve against the measured copy bandwidth) is reported beside it.
  cpu_baseline  the CPU oracle (C port of the same algorithm -- CasADi/Ipopt cannot be installed here) on all
            host cores over a bounded sample of the same workload.

--impl reference times that CPU oracle as the stand-in for the reference's Ipopt path (kind "port").
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "nmpc_solves_per_sec_N40_B65536"
UNIT = "solves/s"
FLOP_STAGE, FLOP_TERM = 2500.0, 500.0  # algorithmic FP64 flop per stage / terminal per IPM iteration


def algorithmic_bytes_per_solve(N: int) -> int:
    """read p = (8N+12) doubles, write z* = (8N+6) doubles + 40 B of scalars."""
    return (8 * N + 12) * 8 + (8 * N + 6) * 8 + 40


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.  In-process NVML calls from a thread
    (a few microseconds each, every 10 ms), started before the untimed settle phase; falls back to one long-lived
    `nvidia-smi -lms 100` process if NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.thread = None
        self.stop = False
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_loop(self, nv, h):
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for n, bit in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import threading

            import pynvml as nv

            nv.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber the devices: address the GPU by the UUID torch reports for it
            h = None
            try:
                import torch

                u = str(torch.cuda.get_device_properties(self.index).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(u if u.startswith("GPU-") else "GPU-" + u)
            except Exception:
                h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                ids = [v for v in vis.split(",") if v.strip().isdigit()]
                h = nv.nvmlDeviceGetHandleByIndex(int(ids[self.index]) if self.index < len(ids) else self.index)
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))
            self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))  # first calls before the timed region
            nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.sm.clear()
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.proc.stdout.readline()  # first sample = process is up; the timed region starts after this
        except Exception:
            self.proc = None

    def mark(self):
        """Forget the samples taken so far (NVML path; the nvidia-smi fallback reports everything it printed)."""
        if self.thread is not None:
            self.sm = []
            self.reasons = set()

    def finish(self) -> dict:
        if self.thread is not None:
            self.stop = True
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        samples = []
        if self.proc is not None:
            time.sleep(0.11)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            samples = [[c.strip() for c in ln.split(",")] for ln in out.splitlines() if ln.strip()]
        sm = [float(s[0]) for s in samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi"}


def workload_config(N: int, B: int, world: int) -> dict:
    return {
        "workload": f"batched NMPC tracking (SURVEY 8(d) config 2 at full size): horizon {N}, {B} perturbed initial "
                    "states per step and GPU around the shipped trajectory (test_cases.json SE(2) families), "
                    "per-problem reference windows, MPCTrackingControl preset (tol 1e-8), cold start",
        "horizon": N, "batch_per_gpu": B, "tol": 1e-8,
        "sharding": f"scenario-parallel x{world}, NCCL gather of u0/status only",
        "cache": f"inputs+outputs ({(B * (16 * N + 18) * 8) >> 20} MiB) and solver scratch (>1 GiB) exceed the 126 MB L2; no flush needed",
    }


def make_batch(cfg, B: int, rank: int):
    from car_trailer_mpc_b200 import problem as pb

    return pb.make_scenarios(cfg, B, seed=20251018 + 7919 * rank)


def run_reference(args, rank: int, world: int) -> None:
    """CPU arm: the oracle port on all host threads, bounded sample per step. Rank 0 only."""
    if rank != 0:
        return
    from car_trailer_mpc_b200 import tracking_preset
    from oracle import oracle

    cores = os.cpu_count() or 1
    cfg = tracking_preset(args.horizon)
    cfg.max_iter = 200
    sample = min(args.batch, 8192)
    sc = make_batch(cfg, sample, 0)
    oracle.build()
    times = []
    iters = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, nthreads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
        iters = r["iters"]
    ms = 1e3 * float(np.mean(times))
    value = sample / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args.horizon, args.batch, args.gpus),
                       cpu_sample=f"each CPU step solves the first {sample} scenarios of the batch"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} scenarios per step x {args.steps} steps, CPU oracle (C, pthreads); "
                                   "CasADi/Ipopt (the reference's solver) is not installable in this image"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mean_iters": float(np.mean(iters)),
    }
    print(json.dumps(line), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="scenarios per step and GPU")
    ap.add_argument("--horizon", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--settle", type=float, default=1.0,
                    help="seconds of untimed solves before the W warm-up steps (clocks and power state settle under the very "
                         "load that is timed)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the N = 100 and OBCA legs of the secondary block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from car_trailer_mpc_b200 import BatchSolver, tracking_preset
    from car_trailer_mpc_b200 import problem as pb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the solver has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_bound = False
    if world > 1:
        # one rank per GPU: run on the CPUs next to that GPU so that the pinned staging buffers of the e2e leg are
        # first-touched on its NUMA node (8 ranks x 350 MB per step otherwise cross the socket interconnect)
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa_bound = True
        except Exception as exc:  # best effort; the measurement is still valid without it
            print(f"[bench] rank {rank}: CPU affinity not set ({exc})", file=sys.stderr)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, B = args.horizon, args.batch
    cfg = tracking_preset(N)
    cfg.max_iter = 200
    sc = make_batch(cfg, B, rank)
    solver = BatchSolver(cfg, local_rank)
    fp64_peak_gflops = solver.measure_fp64_peak()

    # ---- device-resident inputs (value) and pinned host buffers (e2e)
    x_d = torch.from_numpy(sc.x_init).to(dev)
    xs_d = torch.from_numpy(sc.ref_states).to(dev)
    us_d = torch.from_numpy(sc.ref_inputs).to(dev)
    x_h = torch.from_numpy(sc.x_init).pin_memory()
    xs_h = torch.from_numpy(sc.ref_states).pin_memory()
    us_h = torch.from_numpy(sc.ref_inputs).pin_memory()
    nz = 8 * N + 6
    gather_u0 = torch.empty((world * B, 2), dtype=torch.float64, device=dev) if world > 1 else None
    gather_st = torch.empty(world * B, dtype=torch.int32, device=dev) if world > 1 else None

    def step_resident():
        r = solver.solve(x_d, xs_d, us_d)
        if world > 1:  # NCCL only gathers per-scenario results (no exchange inside the solve)
            dist.all_gather_into_tensor(gather_u0, r["u0"])
            dist.all_gather_into_tensor(gather_st, r["status"])
        return r

    # e2e through the library's own host-pointer path: page-locked host buffers in, page-locked host buffers out, three
    # output sets because up to three solves are in flight (TTMPC_FLAG_ASYNC_HOST); device-side timing by the library.
    def pinned(shape, dtype=torch.float64):
        return torch.empty(shape, dtype=dtype).pin_memory()

    out_full = [dict(z=pinned((B, nz)), u0=pinned((B, 2)), obj=pinned((B,)), kkt=pinned((B, 3)), iters=pinned((B,), torch.int32),
                     status=pinned((B,), torch.int32)) for _ in range(3)]
    out_compact = [dict(u0=pinned((B, 2)), iters=pinned((B,), torch.int32), status=pinned((B,), torch.int32)) for _ in range(3)]
    k_h = torch.from_numpy(sc.k_index).pin_memory()
    ti_h = torch.from_numpy(sc.traj_index).pin_memory()
    ts_h = torch.from_numpy(sc.traj_states).pin_memory()
    tu_h = torch.from_numpy(sc.traj_inputs).pin_memory()

    def run_e2e(steps: int, compact: bool, pipelined: bool = True) -> float:
        """Device-side ms for `steps` host-buffer solves (first copy-in to last copy-out)."""
        total = 0.0
        for i in range(steps):
            if compact:
                solver.solve_shared(x_h, k_h, ts_h, tu_h, traj_index=ti_h, want_z=False, want_kkt=False, host_async=True,
                                    out=out_compact[i % 3])
            else:
                solver.solve(x_h, xs_h, us_h, host_async=True, out=out_full[i % 3])
            if not pipelined:
                total += solver.sync()
        return total + (solver.sync() if pipelined else 0.0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        # The warm-up holds its result across the next call exactly like the timed loop does (`r = fn()`): the solve
        # returns freshly allocated output tensors, so the loop needs TWO output sets in torch's caching allocator, and
        # when the second one was first asked for inside the timed region (step index 1) its cudaMalloc -- a device
        # synchronisation of 10..100 ms -- landed there (profiles/r2_bench_1gpu_outlier_step.json).
        r = None
        for _ in range(warmup):
            r = fn()
        barrier()
        l_before = solver.launch_count()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            r = fn()
            ev[i + 1].record()
        barrier()
        per = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(steps)])
        total = ev[0].elapsed_time(ev[steps])
        return total, per, r, solver.launch_count() - l_before

    sampler = ClockSampler(local_rank)
    sampler.start()  # before the settle phase: the sampler's own start-up (first NVML calls) is not inside the timed region
    r_keep = None
    for _ in range(max(0, int(round(args.settle * 150.0)))):  # untimed, ~6.5 ms each: clocks and power state settle under this
        r_keep = step_resident()                              # very load; a COUNT, not a time: every rank runs the same collectives
    torch.cuda.synchronize()
    sampler.mark()   # report the samples from here on: warm-up steps + timed region
    total_ms, per_ms, r, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.finish()
    lanes_headline = solver.last_solve_lanes()
    # the pipelined legs time exactly the K steps of the device-resident region (first copy-in to last copy-out: one
    # copy-in and one copy-out are not overlapped by anything, which is why E approaches `value` from below as K grows);
    # the one-solve-in-flight leg is a per-step figure and keeps a bounded count
    e2e_steps = max(4, args.steps)
    e2e_serial_steps = max(4, min(args.steps, 12))
    run_e2e(3, False)
    barrier()
    e2e_serial_ms = run_e2e(e2e_serial_steps, False, pipelined=False)   # one solve in flight: copy-in, solve, copy-out back to back
    barrier()
    e2e_total_ms = run_e2e(e2e_steps, False)
    barrier()
    run_e2e(3, True)
    barrier()
    e2e_compact_ms = run_e2e(e2e_steps, True)
    barrier()
    assert torch.equal(out_compact[(e2e_steps - 1) % 3]["u0"], out_full[(e2e_steps - 1) % 3]["u0"]), "compact != full contract"

    # p99 of the step latency over >= 200 consecutive steps; the timed region above stays K steps
    lat_steps = max(0, 200 - args.steps)
    if lat_steps:
        _, per_extra, _, _ = timed(step_resident, lat_steps, 0)
        per_lat = np.concatenate([per_ms, per_extra])
    else:
        per_lat = per_ms

    # ---- strong scaling + sharding invariance: the SAME 65 536 scenarios (rank 0's batch of the N = 1 run) split over the ranks
    sc0 = sc if rank == 0 else make_batch(cfg, B, 0)
    lo, hi = rank * B // world, (rank + 1) * B // world
    xg = [torch.from_numpy(np.ascontiguousarray(a[lo:hi])).to(dev) for a in (sc0.x_init, sc0.ref_states, sc0.ref_inputs)]
    u0_all = torch.empty((B, 2), dtype=torch.float64, device=dev) if world > 1 else None

    def step_strong():
        r_ = solver.solve(*xg, want_z=False)
        if world > 1:
            dist.all_gather_into_tensor(u0_all, r_["u0"])
        return r_

    strong_ms, _, r_strong, _ = timed(step_strong, max(4, min(args.steps, 10)), 2)
    strong_ms /= max(4, min(args.steps, 10))
    strong_lanes = solver.last_solve_lanes()
    # checksum with ONE kernel flavour on every rank count (the flavours differ in the last bits: other summation order)
    os.environ["TTMPC_KERNEL"], os.environ["TTMPC_TEAM_LANES"] = "team", "16"
    r_ck = solver.solve(*xg, want_z=False)
    os.environ.pop("TTMPC_KERNEL"), os.environ.pop("TTMPC_TEAM_LANES")
    u0_ck = r_ck["u0"]
    if world > 1:
        dist.all_gather_into_tensor(u0_all, u0_ck)
        u0_ck = u0_all
    torch.cuda.synchronize()
    import hashlib
    checksum = hashlib.sha256(u0_ck.cpu().numpy().tobytes()).hexdigest()[:16]

    # ---- secondary configurations (rank 0 only; each a handful of launches)
    secondary = {}
    if rank == 0:
        def quick(solver_, args_, steps=5, **kw):
            rr = None
            for _ in range(2):  # results held across the next call, as in the timed loop (see timed())
                rr = solver_.solve(*args_, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                rr = solver_.solve(*args_, **kw)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps, rr

        for Bs in (4096, 1):
            ms_, rr = quick(solver, (x_d[:Bs].contiguous(), xs_d[:Bs].contiguous(), us_d[:Bs].contiguous()))
            secondary[f"B{Bs}_N{N}"] = {"ms_per_step": ms_, "solves_per_s": Bs / ms_ * 1e3, "lanes_per_problem": solver.last_solve_lanes(),
                                         "mean_iters": float(rr["iters"].float().mean())}
        if not args.no_secondary:
            cfg100 = tracking_preset(100)
            cfg100.max_iter = 200
            s100 = BatchSolver(cfg100, local_rank)
            sc100 = pb.make_scenarios(cfg100, 8192, seed=1100)
            a100 = tuple(torch.from_numpy(a).to(dev) for a in (sc100.x_init, sc100.ref_states, sc100.ref_inputs))
            ms_, rr = quick(s100, a100, steps=3, want_z=False)
            secondary["B8192_N100"] = {"ms_per_step": ms_, "solves_per_s": 8192 / ms_ * 1e3, "lanes_per_problem": s100.last_solve_lanes(),
                                       "mean_iters": float(rr["iters"].float().mean())}
            s100.close()
            from car_trailer_mpc_b200.config import Obstacles, parking_lot_obstacles

            cfg50 = tracking_preset(50)
            cfg50.max_iter = 300
            s50 = BatchSolver(cfg50, local_rank)
            S_, U_ = pb.load_reference_trajectory(dt=cfg50.dt)
            rng = np.random.default_rng(4)
            ko = rng.integers(0, 341, size=2048).astype(np.int32)
            xo = S_[ko] + rng.normal(0.0, 0.002, size=(2048, 6))
            obs = Obstacles.from_list(parking_lot_obstacles())
            ao = (obs, torch.from_numpy(xo).to(dev), torch.from_numpy(ko).to(dev), torch.from_numpy(S_).to(dev), torch.from_numpy(U_).to(dev))
            s50.solve_obca_shared(*ao, want_z=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ro = s50.solve_obca_shared(*ao, want_z=False)
            e1.record()
            torch.cuda.synchronize()
            sto = ro["status"].cpu().numpy()
            secondary["obca_B2048_N50"] = {"ms_per_step": e0.elapsed_time(e1), "solves_per_s": 2048 / e0.elapsed_time(e1) * 1e3,
                                           "frac_converged": float((sto <= 1).mean()), "mean_iters": float(ro["iters"].float().mean()),
                                           "workload": "11 rectangles of obstacles.json, window starts 0..340, sigma 0.002"}
            # the single solve of the MPCTrackingControlObs shim (how the reference itself uses this controller, once per
            # 50 ms control period): one thread-block cluster per problem
            a1 = (obs, ao[1][:1].contiguous(), ao[2][:1].contiguous(), ao[3], ao[4])
            s50.solve_obca_shared(*a1, want_z=False)
            torch.cuda.synchronize()
            n0 = dict(s50.kernel_launches())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                r1 = s50.solve_obca_shared(*a1, want_z=False)
            e1.record()
            torch.cuda.synchronize()
            n1 = s50.kernel_launches()
            secondary["obca_B1_N50"] = {"ms_per_step": e0.elapsed_time(e1) / 3, "iters": int(r1["iters"][0]), "status": int(r1["status"][0]),
                                        "kernel": [k for k in n1 if n1[k] > n0.get(k, 0)]}
            s50.close()

    iters = r["iters"].cpu().numpy()
    status = r["status"].cpu().numpy()
    kkt = r["kkt"].cpu().numpy()

    # max over ranks (device time)
    t = torch.tensor([total_ms, e2e_total_ms, float(np.percentile(per_lat, 99)), float(np.percentile(per_lat, 50)), e2e_compact_ms,
                      e2e_serial_ms, strong_ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(iters.sum()), float((status <= 1).sum()), float(B)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms, e2e_total_ms, p99_ms, p50_ms, e2e_compact_ms, e2e_serial_ms, strong_ms = [float(v) for v in t.cpu()]
    iters_sum, ok_sum, b_sum = [float(v) for v in agg.cpu()]

    if rank == 0:
        ms_per_step = total_ms / args.steps
        solves_per_step = world * B
        value = solves_per_step / (ms_per_step * 1e-3)
        e2e_value = solves_per_step / (e2e_total_ms / e2e_steps * 1e-3)
        mean_iters = iters_sum / b_sum
        # roofline of the dominant kernel (per GPU, rank 0's launch times; one launch per step)
        kernel_ms = float(np.mean(per_ms))
        flop_per_launch = B * mean_iters * (FLOP_STAGE * N + FLOP_TERM)
        achieved_tf = flop_per_launch / (kernel_ms * 1e-3) * 1e-12
        peak_tf = fp64_peak_gflops * 1e-3
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        bytes_per_launch = B * algorithmic_bytes_per_solve(N)
        hbm_achieved = bytes_per_launch / (kernel_ms * 1e-3) * 1e-9
        traffic = traffic_src = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "solve_kernel_traffic.json")))
            traffic, traffic_src = tj["dram_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "settle_s": args.settle,
            "config": workload_config(N, B, world),
            "p50_step_ms": p50_ms, "p99_step_ms": p99_ms, "timed_step_ms": [round(float(v), 3) for v in per_ms], "latency_steps": int(len(per_lat)), "control_period_ms": 50.0,
            "solve_kernel": "ttmpc_team_kernel" if lanes_headline else "ttmpc_solve_kernel", "lanes_per_problem": lanes_headline,
            "mean_iters": mean_iters, "max_iters": int(iters.max()), "frac_success": ok_sum / b_sum,
            "kkt_max": {"dual_inf": float(kkt[status == 0, 0].max()), "constr_viol": float(kkt[status == 0, 1].max()),
                        "compl": float(kkt[status == 0, 2].max())},
            "roofline": {
                "bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "traffic": traffic, "traffic_source": traffic_src,
                "note": "non-tensor FP64 pipe; algorithmic flop = mean_iters*(2500*N+500) per solve (SURVEY 8(d)); peak = "
                        "DFMA microbenchmark measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                "hbm": {"achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_achieved / hbm_peak,
                        "algorithmic_bytes_per_solve": algorithmic_bytes_per_solve(N), "peak_source": hbm_src},
            },
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(x_h.numel() + xs_h.numel() + us_h.numel()) * 8,
                    "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in out_full[0].values())),
                    "ms_per_step": e2e_total_ms / e2e_steps, "steps": e2e_steps,
                    "mode": "ttmpc_solve_batch with TTMPC_FLAG_HOST_POINTERS|TTMPC_FLAG_ASYNC_HOST on page-locked host buffers: the "
                            "library's own three-stream copy-in | solve | copy-out pipeline, up to three solves in flight; timed "
                            "with CUDA events on the library's copy streams (ttmpc_host_pipeline_ms)",
                    "cpu_affinity_to_gpu_numa_node": numa_bound, "serial_ms_per_step": e2e_serial_ms / e2e_serial_steps},
            "e2e_compact": {"value": solves_per_step / (e2e_compact_ms / e2e_steps * 1e-3), "unit": UNIT,
                            "h2d_bytes_per_step": int(x_h.numel()) * 8 + int(k_h.numel() + ti_h.numel()) * 4 + int(ts_h.numel() + tu_h.numel()) * 8,
                            "d2h_bytes_per_step": int(sum(v.numel() * v.element_size() for v in out_compact[0].values())),
                            "ms_per_step": e2e_compact_ms / e2e_steps, "steps": e2e_steps,
                            "mode": "ttmpc_solve_batch_multi, same flags: x_init + window start + trajectory index in, u0 / iterations / "
                                    "status out; identical first controls (checked in this run)"},
            "strong": {"B_total": B, "n_gpus": world, "ms_per_step": strong_ms, "solves_per_s": B / (strong_ms * 1e-3),
                       "lanes_per_problem": strong_lanes,
                       "note": "the N = 1 run's 65 536 scenarios split over the ranks, automatic kernel choice, u0 gathered with NCCL"},
            "u0_checksum_global": checksum,
            "secondary": secondary,
            "gpu_launches": int(launches),
            "kernels": {k: v for k, v in solver.kernel_launches().items() if v},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            from oracle import oracle

            cores = os.cpu_count() or 1
            sample = min(B, 16384)
            oracle.build()
            oracle.solve_batch(cfg, sc.x_init[:256], sc.ref_states[:256], sc.ref_inputs[:256], nthreads=cores)
            t0 = time.perf_counter()
            ro = oracle.solve_batch(cfg, sc.x_init[:sample], sc.ref_states[:sample], sc.ref_inputs[:sample], nthreads=cores)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {sample} scenarios of the batch, CPU oracle (C, pthreads, all host threads); "
                                              "stand-in for CasADi/Ipopt, which is not installable in this image",
                                    "mean_iters": float(ro["iters"].mean())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE line (the JSON record): libraries that print to fd 1 (NCCL's version banner, for one)
    # are sent to stderr for the duration of the run, and the record is written to the real stdout at the end.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real_stdout, "w", buffering=1)
    main()
