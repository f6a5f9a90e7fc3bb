"""Known answers from an independent ALGORITHM: tests/golden/slsqp_*.npz hold SciPy-SLSQP solutions (active-set SQP) of the
literal reference NLP (x_0 a bounded variable pinned by an equality, trajectory_planning.py:28-60; tests/nlp_numpy.py),
64 problems per configuration, made by tools/make_golden_slsqp.py without the oracle or the kernels.  The oracle, the host
builds of both kernel cores and (tests/test_gpu_team.py) the GPU must reach SLSQP's minimiser -- so an error in the
shared reading of Ipopt's rules could not hide behind oracle == kernel."""
import glob
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import emu  # noqa: E402
import make_golden_slsqp as mk  # noqa: E402
import nlp_numpy as nlp  # noqa: E402

from oracle import oracle  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(GOLD, "slsqp_*.npz")))


def load(path):
    g = np.load(path)
    name = os.path.basename(path)[len("slsqp_"):-len(".npz")]
    return name, mk.config_of(mk.CONFIGS[name]), g


def check_against_slsqp(cfg, g, r, label):
    ok = g["success"] == 1
    assert ok.sum() >= 60, f"{label}: SLSQP solved only {ok.sum()} of {len(ok)}"
    assert (np.asarray(r["status"])[ok] == 0).all(), f"{label}: {np.bincount(np.asarray(r['status'])[ok])}"
    N = cfg.horizon
    dj = np.abs(np.asarray(r["obj"]) - g["obj"]) / np.maximum(1.0, np.abs(g["obj"]))
    du0 = np.abs(np.asarray(r["u0"]) - g["z"][:, 6:8]).max(1)
    dz = np.abs(np.asarray(r["z"]) - g["z"]).max(1)
    assert dj[ok].max() < 1e-6, f"{label}: objective {dj[ok].max()}"
    assert du0[ok].max() < 1e-4, f"{label}: u0 {du0[ok].max()}"          # the north star's tolerance on first controls
    assert dz[ok].max() < 2e-4, f"{label}: z {dz[ok].max()}"
    assert 8 * N + 6 == g["z"].shape[1]


def test_fixtures_cover_the_promised_configurations():
    names = {os.path.basename(f)[len("slsqp_"):-len(".npz")] for f in FILES}
    assert names == set(mk.CONFIGS), names
    for f in FILES:
        g = np.load(f)
        assert g["z"].shape[0] == 64 and str(g["generator"]) == "tools/make_golden_slsqp.py" and int(g["seed"]) > 0


@pytest.mark.parametrize("path", FILES, ids=os.path.basename)
def test_oracle_reaches_slsqp_minimiser(path):
    name, cfg, g = load(path)
    r = oracle.solve_batch(cfg, g["x_init"], g["ref_states"], g["ref_inputs"], nthreads=4)
    check_against_slsqp(cfg, g, r, name)


@pytest.mark.parametrize("path", FILES, ids=os.path.basename)
def test_kernel_cores_reach_slsqp_minimiser(path):
    """Host builds of the lane-per-problem core and of the warp-cooperative team core."""
    name, cfg, g = load(path)
    n = 16 if cfg.horizon >= 100 else 64
    sub = {k: g[k][:n] for k in ("x_init", "ref_states", "ref_inputs", "z", "obj", "success")}
    sub["success"] = np.where(np.arange(n) < n, sub["success"], 0)
    for label, r in (("lane", emu.solve_batch(cfg, sub["x_init"], sub["ref_states"], sub["ref_inputs"])),
                     ("team", emu.team_solve_batch(cfg, sub["x_init"], sub["ref_states"], sub["ref_inputs"],
                                                   lanes=32 if cfg.horizon >= 100 else 16))):
        ok = sub["success"] == 1
        assert (r["status"][ok] == 0).all(), (name, label)
        assert (np.abs(r["obj"] - sub["obj"]) / np.maximum(1.0, np.abs(sub["obj"])))[ok].max() < 1e-6, (name, label)
        assert np.abs(r["u0"] - sub["z"][:, 6:8])[ok].max() < 1e-4, (name, label)


def test_fixture_is_reproducible():
    """Re-run SLSQP live on two problems of the config-2 fixture: the committed answers are what the generator gives."""
    name, cfg, g = load(os.path.join(GOLD, "slsqp_config2_narrow_N40.npz"))
    for i in (0, 33):
        z, obj, success, _ = mk.solve_one((mk.CONFIGS[name], g["x_init"][i], g["ref_states"][i], g["ref_inputs"][i]))
        assert success == int(g["success"][i])
        assert abs(obj - g["obj"][i]) < 1e-9 and np.abs(z - g["z"][i]).max() < 1e-6
        stat, viol, bviol, neg = nlp.kkt_certificate(cfg, z, g["x_init"][i], g["ref_states"][i], g["ref_inputs"][i], active_tol=1e-7)
        assert viol < 1e-8 and bviol < 1e-12 and neg > -1e-6
