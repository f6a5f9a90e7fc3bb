"""Multi-rank host logic on CPU: world_size-2 gloo, the "device" is the CPU oracle (no GPU needed).
Sharding must not change answers: the gathered result equals the single-process result bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from car_trailer_mpc_b200 import problem as pb
    from car_trailer_mpc_b200 import sharding, tracking_preset
    from oracle import oracle

    cfg = tracking_preset(20)
    sc = pb.make_scenarios(cfg, B, seed=5)
    got = sharding.solve_sharded(lambda x, xs, us: oracle.solve_batch(cfg, x, xs, us), sc.x_init, sc.ref_states,
                                 sc.ref_inputs, rank, world)
    lo, hi = sharding.shard_range(B, rank, world)
    red = sharding.reduce_metrics({"solved": hi - lo, "iters_max": int(got["iters"][lo:hi].max()), "lo_min": lo})
    if rank == 0:
        np.savez(out_path, u0=got["u0"].numpy(), status=got["status"].numpy(), iters=got["iters"].numpy(),
                 solved=red["solved"], iters_max=red["iters_max"], lo_min=red["lo_min"])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [16, 17])   # even and ragged split
def test_two_rank_gloo_gather_equals_single_process(tmp_path, B):
    sys.path.insert(0, ROOT)
    from car_trailer_mpc_b200 import problem as pb
    from car_trailer_mpc_b200 import tracking_preset
    from oracle import oracle

    port = 29500 + (os.getpid() % 2000) + B
    out = str(tmp_path / "g.npz")
    mp.spawn(_worker, args=(2, port, B, out), nprocs=2, join=True)
    g = np.load(out)
    cfg = tracking_preset(20)
    sc = pb.make_scenarios(cfg, B, seed=5)
    ref = oracle.solve_batch(cfg, sc.x_init, sc.ref_states, sc.ref_inputs)
    assert np.array_equal(g["u0"], ref["u0"]) and np.array_equal(g["status"], ref["status"]) and np.array_equal(g["iters"], ref["iters"])
    assert g["solved"] == B and g["iters_max"] == ref["iters"].max() and g["lo_min"] == 0


def test_shard_ranges_partition_everything():
    from car_trailer_mpc_b200 import sharding
    for total in (0, 1, 7, 8, 65536, 1000003):
        for world in (1, 2, 3, 8):
            rs = [sharding.shard_range(total, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == total
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
            assert sharding.shard_sizes(total, world) == [b - a for a, b in rs]
