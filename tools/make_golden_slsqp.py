"""Known answers from an INDEPENDENT algorithm: SciPy SLSQP (active-set SQP) on the literal NLP of the reference
(tests/nlp_numpy.py: x_0 a bounded decision variable pinned by an equality, exactly as trajectory_planning.py:28-60
builds it), 64 problems per configuration.  Neither the oracle nor the kernels are involved in producing these files.

  python tools/make_golden_slsqp.py [names...]      -> tests/golden/slsqp_<name>.npz

Configurations: config 2 narrow and wide (N = 40), N = 10, N = 100, and the TruckTrailerNMPC preset (bounds / weights of
simulation_nmpc.py:124-148) solved to optimality.  Each file holds the inputs, SLSQP's z and objective, its exit status
and the provenance (seed, scipy version, git commit).  SLSQP runs from the reference window (the controllers' cold
start); a problem on which it reports failure is kept with success = 0 and skipped by the tests."""
import multiprocessing as mp
import os
import subprocess
import sys
import time

import numpy as np
import scipy
from scipy.optimize import minimize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nlp_numpy as nlp  # noqa: E402
from car_trailer_mpc_b200 import nmpc_preset, tracking_preset  # noqa: E402
from car_trailer_mpc_b200 import problem as pb  # noqa: E402

COUNT = 64
CONFIGS = {
    "config2_narrow_N40": dict(preset="tracking", N=40, sigma="narrow", seed=4101),
    "config2_wide_N40": dict(preset="tracking", N=40, sigma="wide", seed=4102),
    "tracking_N10": dict(preset="tracking", N=10, sigma="narrow", seed=4103),
    "tracking_N100": dict(preset="tracking", N=100, sigma="narrow", seed=4104),
    "nmpc_bounds_N30": dict(preset="nmpc", N=30, sigma="narrow", seed=4105),
}


def config_of(spec):
    cfg = tracking_preset(spec["N"]) if spec["preset"] == "tracking" else nmpc_preset(spec["N"])
    cfg.tol, cfg.acceptable_tol, cfg.acceptable_iter, cfg.max_iter = 1e-8, 1e-6, 15, 300
    return cfg


def solve_one(job):
    spec, x0, xs, us = job
    cfg = config_of(spec)
    lb, ub = nlp.bounds(cfg)
    bnds = [(None if not np.isfinite(l) else l, None if not np.isfinite(u) else u) for l, u in zip(lb, ub)]
    z0 = pb.pack_z(xs, us)
    z0[:6] = x0  # start on the equality x_0 = x_init
    res = minimize(lambda z: nlp.cost(cfg, z, xs, us), z0, jac=lambda z: nlp.cost_grad(cfg, z, xs, us), method="SLSQP", bounds=bnds,
                   constraints=[{"type": "eq", "fun": lambda z: nlp.constraints(cfg, z, x0), "jac": lambda z: nlp.constraints_jac(cfg, z, x0)}],
                   options={"ftol": 1e-12, "maxiter": 400})
    viol = float(np.abs(nlp.constraints(cfg, res.x, x0)).max())
    return res.x, float(res.fun), int(bool(res.success) and viol < 1e-8), int(res.nit)


def main():
    names = sys.argv[1:] or list(CONFIGS)
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    for name in names:
        spec = CONFIGS[name]
        cfg = config_of(spec)
        sig = pb.SIGMA_NARROW if spec["sigma"] == "narrow" else pb.SIGMA_WIDE
        sc = pb.make_scenarios(cfg, COUNT, seed=spec["seed"], sigma=sig)
        t0 = time.time()
        with mp.Pool(min(8, os.cpu_count() or 1)) as pool:
            out = pool.map(solve_one, [(spec, sc.x_init[i], sc.ref_states[i], sc.ref_inputs[i]) for i in range(COUNT)])
        z = np.stack([o[0] for o in out])
        np.savez_compressed(
            os.path.join(ROOT, "tests", "golden", f"slsqp_{name}.npz"), x_init=sc.x_init, ref_states=sc.ref_states, ref_inputs=sc.ref_inputs,
            z=z, obj=np.array([o[1] for o in out]), success=np.array([o[2] for o in out], dtype=np.int32),
            nit=np.array([o[3] for o in out], dtype=np.int32), preset=spec["preset"], horizon=spec["N"], sigma=spec["sigma"],
            seed=spec["seed"], scipy_version=scipy.__version__, generator="tools/make_golden_slsqp.py", git_commit=commit)
        print(f"{name}: {sum(o[2] for o in out)}/{COUNT} solved, {time.time() - t0:.0f} s", flush=True)


if __name__ == "__main__":
    main()
