"""Shared helpers of the obstacle-aware (OBCA) tests."""
import os

import numpy as np

from car_trailer_mpc_b200 import tracking_preset
from car_trailer_mpc_b200.config import Obstacles

GOLD_OBCA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obca_cases.npz")
# states and inputs of two converged solves of the same NLP agree far better than the north-star tolerances (the last
# iterations at the final barrier parameter contract quadratically); the north-star numbers are the hard bar
Z_TOL = 1e-6


GOLD_OBCA_FULL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "obca_cases_full.npz")


def golden_cases(path=GOLD_OBCA):
    g = np.load(path)
    names = sorted({k.split("/")[0] for k in g.files if "/" in k})
    out = []
    for n in names:
        c = {k.split("/")[1]: g[k] for k in g.files if k.startswith(n + "/")}
        c["name"] = n
        out.append(c)
    return out


def case_problem(c):
    N = int(c["horizon"])
    cfg = tracking_preset(N)
    obs = Obstacles.from_list([tuple(r) for r in c["rects"]])
    return cfg, obs


def split_z(z, N):
    z = np.asarray(z)
    xs = np.stack([z[..., 8 * k:8 * k + 6] for k in range(N + 1)], axis=-2)
    us = np.stack([z[..., 8 * k + 6:8 * k + 8] for k in range(N)], axis=-2)
    return xs, us


# the three kernels of the obstacle-aware path and the environment that forces each of them (ttmpc.cu: obca_device)
OBCA_FLAVOURS = ["warp_per_problem", "cta_per_problem", "cluster_per_problem"]
OBCA_KERNEL = {"warp_per_problem": "ttmpc_obca_kernel", "cta_per_problem": "ttmpc_obca_wide_kernel",
               "cluster_per_problem": "ttmpc_obca_cluster_kernel"}


def force_obca_kernel(monkeypatch, flavour, cluster="8"):
    monkeypatch.delenv("TTMPC_OBCA_WIDE_MAX", raising=False)
    monkeypatch.delenv("TTMPC_OBCA_CLUSTER", raising=False)
    if flavour == "warp_per_problem":
        monkeypatch.setenv("TTMPC_OBCA_WIDE_MAX", "0")
    elif flavour == "cta_per_problem":
        monkeypatch.setenv("TTMPC_OBCA_CLUSTER", "0")
    else:
        monkeypatch.setenv("TTMPC_OBCA_CLUSTER", cluster)
    return OBCA_KERNEL[flavour]
