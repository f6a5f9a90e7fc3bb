"""Scenario-parallel sharding over the GPUs of one box.

Every MPC problem / closed-loop episode is independent, so the path shards by scenario index with NO data-path
collective: rank r of G owns the contiguous block ``[r*B/G, (r+1)*B/G)``; after a step (or an
episode) ``torch.distributed`` only gathers the per-scenario result rows and reduces a handful of metric
accumulators -- NCCL on GPUs, gloo in the CPU tests (world_size 2).
"""
from __future__ import annotations

import numpy as np


def shard_range(total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition; the first ``total % world`` ranks get one extra scenario."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int) -> list[int]:
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_rows(local, total: int, group=None):
    """All-gather per-scenario rows ``[n_local, ...]`` (torch tensor, any backend) into ``[total, ...]`` in global
    scenario order.  Ragged shards are padded to the largest shard for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(total, world)
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    parts = [out[r * m : r * m + sizes[r]] for r in range(world)]
    return torch.cat(parts, dim=0)


def reduce_metrics(values: dict, group=None) -> dict:
    """All-reduce a dict of scalar accumulators: keys ending in ``_max`` use MAX, ``_min`` MIN, everything else SUM."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return dict(values)
    keys = sorted(values)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    out = {}
    for op, sel in ((dist.ReduceOp.MAX, lambda k: k.endswith("_max")), (dist.ReduceOp.MIN, lambda k: k.endswith("_min")),
                    (dist.ReduceOp.SUM, lambda k: not (k.endswith("_max") or k.endswith("_min")))):
        ks = [k for k in keys if sel(k)]
        if not ks:
            continue
        t = torch.tensor([float(values[k]) for k in ks], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op, group=group)
        out.update({k: float(v) for k, v in zip(ks, t.cpu())})
    return out


def solve_sharded(solve_fn, x_init: np.ndarray, ref_states: np.ndarray, ref_inputs: np.ndarray, rank: int, world: int):
    """Solve this rank's block of a global batch with ``solve_fn(x, xs, us) -> dict`` (a :class:`BatchSolver` method on
    a GPU, or the CPU oracle in the gloo tests) and gather ``u0``/``status``/``iters``/``obj`` in global order."""
    import torch

    B = x_init.shape[0]
    lo, hi = shard_range(B, rank, world)
    r = solve_fn(x_init[lo:hi], ref_states[lo:hi], ref_inputs[lo:hi])
    out = {}
    for key in ("u0", "status", "iters", "obj"):
        v = r[key]
        t = v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v))
        out[key] = gather_rows(t, B)
    return out
