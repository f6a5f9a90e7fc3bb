"""GPU tests of round 2's additions, all through the C ABI: both flavours of the plain solve (the lane-per-problem kernel
and the warp-cooperative team kernel with 8 / 16 / 32 lanes per problem) against the CPU oracle, the dispatch between
them, the compact multi-trajectory contract, the asynchronous host-pointer pipeline, stream ordering, and a
solver-independent KKT certificate computed on the GPU's own output."""
import os
import sys

import numpy as np
import pytest

import nlp_numpy as nlp
from parity import assert_parity

from car_trailer_mpc_b200 import tracking_preset
from car_trailer_mpc_b200 import problem as pb

pytestmark = pytest.mark.gpu

FLAVOURS = {"lane": {"TTMPC_KERNEL": "lane"}, "team8": {"TTMPC_KERNEL": "team", "TTMPC_TEAM_LANES": "8"},
            "team16": {"TTMPC_KERNEL": "team", "TTMPC_TEAM_LANES": "16"}, "team32": {"TTMPC_KERNEL": "team", "TTMPC_TEAM_LANES": "32"}}


def force(monkeypatch, flavour):
    for k in ("TTMPC_KERNEL", "TTMPC_TEAM_LANES"):
        monkeypatch.delenv(k, raising=False)
    for k, v in FLAVOURS.get(flavour, {}).items():
        monkeypatch.setenv(k, v)


def make_solver(cfg):
    from car_trailer_mpc_b200 import BatchSolver
    return BatchSolver(cfg, 0)


@pytest.fixture(scope="module")
def config2():
    from oracle import oracle
    cfg = tracking_preset(40); cfg.max_iter = 200
    out = {}
    for name, sig in (("narrow", pb.SIGMA_NARROW), ("wide", pb.SIGMA_WIDE)):
        sc = pb.make_scenarios(cfg, 4096, sigma=sig)
        out[name] = (sc, oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs, nthreads=os.cpu_count() or 1))
    return cfg, out


@pytest.mark.parametrize("flavour", list(FLAVOURS))
@pytest.mark.parametrize("sigma_name", ["narrow", "wide"])
def test_every_flavour_matches_the_oracle_on_config2(flavour, sigma_name, config2, monkeypatch):
    cfg, data = config2
    sc, ref = data[sigma_name]
    force(monkeypatch, flavour)
    s = make_solver(cfg)
    r = s.solve(sc.x_init, sc.ref_states, sc.ref_inputs)
    assert s.last_solve_lanes(host=True) == {"lane": 0, "team8": 8, "team16": 16, "team32": 32}[flavour]
    assert_parity(cfg, r, ref, sc.x_init, label=f"{flavour}/{sigma_name}")
    assert (r["iters"] == ref["iters"]).mean() > 0.98
    # the reported KKT residuals are honest: recompute the constraint violation and the objective from z
    X, U = pb.unpack_z(r["z"], 40)
    viol = np.abs(pb.dynamics_defect(cfg, X, U)).reshape(4096, -1).max(1)
    assert np.abs(viol - r["kkt"][:, 1]).max() < 1e-12
    assert np.abs(pb.objective(cfg, X, U, sc.ref_states, sc.ref_inputs) - r["obj"]).max() < 1e-10


def test_dispatch_picks_team_for_small_and_lane_for_large_batches(monkeypatch):
    import torch
    force(monkeypatch, None)
    cfg = tracking_preset(40); cfg.max_iter = 200
    s = make_solver(cfg)
    dev = torch.device("cuda:0")
    sc = pb.make_scenarios(cfg, 65536)
    x, xs, us = (torch.from_numpy(a).to(dev) for a in (sc.x_init, sc.ref_states, sc.ref_inputs))
    picked = {}
    for B in (1, 100, 4096, 65536):
        s.solve(x[:B].contiguous(), xs[:B].contiguous(), us[:B].contiguous(), want_z=False)
        picked[B] = s.last_solve_lanes()
    torch.cuda.synchronize()
    assert picked[1] == 32 and picked[100] in (16, 32) and picked[4096] in (8, 16) and picked[65536] == 0, picked


@pytest.mark.parametrize("flavour", ["lane", "team16"])
def test_results_do_not_depend_on_batch_composition(flavour, config2, monkeypatch):
    """Problems are independent: any sub-batch, in any order, gives bit-identical results (the basis of sharding
    invariance -- 1 GPU vs N GPUs -- and of the hardest-first scheduling order)."""
    cfg, data = config2
    sc, _ = data["wide"]
    force(monkeypatch, flavour)
    s = make_solver(cfg)
    full = s.solve(sc.x_init, sc.ref_states, sc.ref_inputs)
    rng = np.random.default_rng(3)
    pick = rng.permutation(4096)[:777]
    part = s.solve(sc.x_init[pick], sc.ref_states[pick], sc.ref_inputs[pick])
    monkeypatch.setenv("TTMPC_NO_ORDER", "1")
    plain = s.solve(sc.x_init, sc.ref_states, sc.ref_inputs)
    for key in ("z", "u0", "obj", "kkt", "iters", "status"):
        assert np.array_equal(part[key], full[key][pick]), key
        assert np.array_equal(plain[key], full[key]), key


@pytest.mark.parametrize("flavour", ["lane", "team16"])
def test_compact_multi_trajectory_contract_equals_window_contract(flavour, monkeypatch):
    """ttmpc_solve_batch_multi (x_init + window start + trajectory index in, u0 / status out) == per-problem windows."""
    import torch
    force(monkeypatch, flavour)
    cfg = tracking_preset(40); cfg.max_iter = 200
    sc = pb.make_scenarios(cfg, 3000, seed=11)
    assert sc.traj_states.shape[0] > 1 and len(np.unique(sc.traj_index)) > 1
    s = make_solver(cfg)
    a = s.solve(sc.x_init, sc.ref_states, sc.ref_inputs)
    b = s.solve_shared(sc.x_init, sc.k_index, sc.traj_states, sc.traj_inputs, traj_index=sc.traj_index)
    for key in ("z", "u0", "obj", "kkt", "iters", "status"):
        assert np.array_equal(a[key], b[key]), key
    c = s.solve_shared(sc.x_init, sc.k_index, sc.traj_states, sc.traj_inputs, traj_index=sc.traj_index, want_z=False, want_kkt=False)
    assert c["z"] is None and c["kkt"] is None and c["obj"] is None
    assert np.array_equal(c["u0"], a["u0"]) and np.array_equal(c["status"], a["status"])
    dev = torch.device("cuda:0")
    d = s.solve_shared(*(torch.from_numpy(v).to(dev) for v in (sc.x_init, sc.k_index, sc.traj_states, sc.traj_inputs)),
                       traj_index=torch.from_numpy(sc.traj_index).to(dev), want_z=False)
    assert np.array_equal(d["u0"].cpu().numpy(), a["u0"])
    with pytest.raises(ValueError):
        s.solve_shared(sc.x_init, sc.k_index, sc.traj_states[0], sc.traj_inputs[0], traj_index=sc.traj_index)


def test_async_host_pipeline_equals_synchronous_calls(monkeypatch):
    """TTMPC_FLAG_ASYNC_HOST: several host-pointer solves in flight (copy-in | solve | copy-out overlap across calls)
    deliver exactly what the synchronous calls deliver, for page-locked and for pageable buffers."""
    import torch
    force(monkeypatch, None)
    cfg = tracking_preset(40); cfg.max_iter = 200
    s = make_solver(cfg)
    batches = [pb.make_scenarios(cfg, 3000 + 500 * i, seed=100 + i) for i in range(5)]
    want = [s.solve(sc.x_init, sc.ref_states, sc.ref_inputs) for sc in batches]
    for pinned in (True, False):
        pin = (lambda a: torch.from_numpy(a).pin_memory()) if pinned else (lambda a: a.copy())
        ins = [tuple(pin(a) for a in (sc.x_init, sc.ref_states, sc.ref_inputs)) for sc in batches]
        got = [s.solve(*i3, host_async=True) for i3 in ins]
        ms = s.sync()
        assert ms > 0.0
        for w, g in zip(want, got):
            for key in ("z", "u0", "obj", "kkt", "iters", "status"):
                assert np.array_equal(np.asarray(g[key]), w[key]), (pinned, key)
    # reusing preallocated (page-locked) output buffers, compact contract
    sc = batches[0]
    out = dict(u0=torch.empty((3000, 2), dtype=torch.float64).pin_memory(), iters=torch.empty(3000, dtype=torch.int32).pin_memory(),
               status=torch.empty(3000, dtype=torch.int32).pin_memory())
    r = s.solve_shared(sc.x_init, sc.k_index, sc.traj_states, sc.traj_inputs, traj_index=sc.traj_index, want_z=False,
                       want_kkt=False, host_async=True, out=out)
    s.sync()
    assert r["u0"] is out["u0"] and np.array_equal(out["u0"].numpy(), want[0]["u0"]) and np.array_equal(out["status"].numpy(), want[0]["status"])


def test_calls_on_different_streams_share_a_handle_safely(monkeypatch):
    """Two un-synchronised calls on different streams through one handle (shared scratch / work queue): the library
    orders them on the device."""
    import torch
    cfg = tracking_preset(40); cfg.max_iter = 200
    dev = torch.device("cuda:0")
    sc = pb.make_scenarios(cfg, 40000, seed=5)
    x, xs, us = (torch.from_numpy(a).to(dev) for a in (sc.x_init, sc.ref_states, sc.ref_inputs))
    for flavour in ("lane", "team16"):
        force(monkeypatch, flavour)
        s = make_solver(cfg)
        ref = s.solve(x, xs, us)
        torch.cuda.synchronize()
        sa, sb = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        half = 20000
        with torch.cuda.stream(sa):
            ra = s.solve(x[:half].contiguous(), xs[:half].contiguous(), us[:half].contiguous(), stream=sa.cuda_stream)
        with torch.cuda.stream(sb):
            rb = s.solve(x[half:].contiguous(), xs[half:].contiguous(), us[half:].contiguous(), stream=sb.cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(ra["u0"], ref["u0"][:half]) and torch.equal(rb["u0"], ref["u0"][half:]), flavour
        assert torch.equal(ra["z"], ref["z"][:half]) and torch.equal(rb["z"], ref["z"][half:]), flavour
    assert torch.cuda.current_device() == 0


def test_bad_arguments_raise_instead_of_reaching_the_kernels():
    import torch
    cfg = tracking_preset(40)
    s = make_solver(cfg)
    sc = pb.make_scenarios(cfg, 8)
    dev = torch.device("cuda:0")
    x, xs, us = (torch.from_numpy(a).to(dev) for a in (sc.x_init, sc.ref_states, sc.ref_inputs))
    with pytest.raises(ValueError):
        s.solve(x.float(), xs, us)                      # wrong dtype
    with pytest.raises(ValueError):
        s.solve(x, xs[:, :-1].contiguous(), us)         # wrong shape
    with pytest.raises(ValueError):
        s.solve(x, xs.cpu(), us)                        # wrong device
    with pytest.raises(ValueError):
        s.solve(sc.x_init, sc.ref_states, sc.ref_inputs, q_weights=np.ones((8, 6)))   # weights go together


def test_kkt_certificate_on_the_gpu_output_at_full_batch(monkeypatch):
    """Solver-independent proof on the PRODUCT's output (not the oracle's): for a 4 096-problem sample of the 65 536
    batch of the headline benchmark, stationarity with least-squares multipliers on the active set read off z, the
    equality constraints of trajectory_planning.py:28-36 with x_0 as a variable, the box, and the multiplier signs."""
    import torch
    force(monkeypatch, None)
    cfg = tracking_preset(40); cfg.max_iter = 200
    sc = pb.make_scenarios(cfg, 65536)
    dev = torch.device("cuda:0")
    s = make_solver(cfg)
    r = s.solve(*(torch.from_numpy(a).to(dev) for a in (sc.x_init, sc.ref_states, sc.ref_inputs)))
    torch.cuda.synchronize()
    assert s.last_solve_lanes() == 0                    # the bulk kernel
    z, st = r["z"].cpu().numpy(), r["status"].cpu().numpy()
    assert (st == 0).all()
    sample = np.random.default_rng(1).choice(65536, 4096, replace=False)
    force(monkeypatch, "team16")                        # and the same sample through the team kernel
    rt = s.solve(sc.x_init[sample], sc.ref_states[sample], sc.ref_inputs[sample])
    assert np.abs(rt["z"] - z[sample]).max() < 1e-6
    worst = np.zeros(5)
    for zz in (z[sample], rt["z"]):
        for j, i in enumerate(sample):
            c5 = nlp.kkt_certificate(cfg, zz[j], sc.x_init[i], sc.ref_states[i], sc.ref_inputs[i], active_tol=1e-3, with_compl=True)
            worst = np.maximum(worst, [c5[0], c5[1], c5[2], -c5[3], c5[4]])
    # stationarity, equality violation (tol = 1e-8), bound violation (bound_relax 1e-8 * max(1,|b|)), multiplier sign,
    # complementarity (mu_final ~ 1e-9)
    assert worst[0] < 5e-6 and worst[1] < 1.5e-8 and worst[2] < 2e-8 and worst[3] < 1e-6 and worst[4] < 2e-7, worst


@pytest.mark.parametrize("flavour", ["lane", "team"])
def test_gpu_reaches_the_slsqp_minimiser_on_every_configuration(flavour, monkeypatch):
    """The product against an independent algorithm: tests/golden/slsqp_*.npz (SciPy SLSQP on the literal reference NLP,
    64 problems per configuration, tools/make_golden_slsqp.py) -- not against the oracle."""
    import glob
    from test_slsqp_golden import GOLD, check_against_slsqp, load
    for k in ("TTMPC_KERNEL", "TTMPC_TEAM_LANES"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("TTMPC_KERNEL", flavour)
    files = sorted(glob.glob(os.path.join(GOLD, "slsqp_*.npz")))
    assert len(files) == 5
    for path in files:
        name, cfg, g = load(path)
        s = make_solver(cfg)
        r = s.solve(g["x_init"], g["ref_states"], g["ref_inputs"])
        assert (s.last_solve_lanes(host=True) == 0) == (flavour == "lane")
        check_against_slsqp(cfg, g, r, f"{flavour}/{name}")


def test_bulk_staged_experiment_build_is_bit_identical():
    """The cp.async.bulk staging of the warp tile (StageBulk, -DTTMPC_STAGE_BULK=1; measured slower and not shipped,
    profiles/r2_bulk_stage_ab.txt) must give the shipped kernel's bits: same u0 for 16 384 problems, for any slot
    assignment.  Needs the experiment library (tools/ab/libttmpc_bulk.so, built with
    TTMPC_NVCC_FLAGS=-DTTMPC_STAGE_BULK=1 TTMPC_BUILD_OUT=... python -m car_trailer_mpc_b200.build --force)."""
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "tools", "ab", "libttmpc_bulk.so")
    srcs = [os.path.join(root, "car_trailer_mpc_b200", "csrc", f) for f in ("ttmpc.cu", "ttmpc_core.cuh", "ttmpc_obca.cuh", "ttmpc_team.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(f) for f in srcs + [os.path.join(root, "include", "ttmpc.h")]):
        pytest.skip("experiment library not built (or older than the sources)")
    out = {}
    for name, env in (("shipped", {}), ("bulk", {"TTMPC_LIB": lib})):
        e = dict(os.environ, **env)
        e.pop("TTMPC_LIB", None) if not env else None
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "perm_check.py"), "16384", "40"], env=e, capture_output=True,
                           text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        m = re.search(r"mismatches=(\d+) .* u0_sha=(\w+) lanes=(\d+)", r.stdout)
        assert m, r.stdout
        assert int(m.group(1)) == 0 and int(m.group(3)) == 0  # permutation-invariant, lane kernel
        out[name] = m.group(2)
    assert out["shipped"] == out["bulk"]
