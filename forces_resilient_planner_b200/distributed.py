"""Multi-GPU driver: contiguous sharding of independent problems, one process per GPU.

The NMPC instances never interact (single-vehicle planner), so the solve itself
needs NO collective: rank r owns problems [r*B/G, (r+1)*B/G) and runs the same fused kernel on
them.  The only exchange is the optional end-of-batch collation of results (z, exit flags,
iteration counts) -- one all_gather over NCCL/NVLink, outside the solve.
"""
from __future__ import annotations

import numpy as np

from .workloads import Batch


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous split; the first B % world ranks take one extra problem."""
    base, extra = divmod(B, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(batch: Batch, rank: int, world: int) -> Batch:
    lo, hi = shard_range(batch.B, rank, world)
    return batch.slice(lo, hi)


def all_gather_results(z_local, flag_local, it_local, B_total: int, group=None):
    """Collate per-rank results into full-batch tensors on every rank.

    Inputs are torch tensors on the backend's device (cuda for nccl, cpu for gloo) holding this
    rank's shard; shards may differ by one problem, so they are padded to the largest shard for
    the fixed-size all_gather and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    per = -(-B_total // world)
    n_local = z_local.shape[0]

    def pad(t):
        if t.shape[0] == per:
            return t.contiguous()
        out = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        out[:n_local] = t
        return out

    outs = []
    for t in (z_local, flag_local, it_local):
        buf = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, pad(t), group=group)
        parts = []
        for r in range(world):
            lo, hi = shard_range(B_total, r, world)
            parts.append(buf[r * per:r * per + (hi - lo)])
        outs.append(torch.cat(parts, 0))
    return tuple(outs)


def solve_sharded(batch: Batch, local_solver, device, gather: bool = True, group=None):
    """Solve `batch` across the process group.  `local_solver(shard) -> (z, flag, it)` numpy arrays
    (the CUDA solver in production; tests inject a CPU stand-in to exercise the plumbing on gloo)."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    mine = shard(batch, rank, world)
    z, flag, it = local_solver(mine)
    if not gather:
        return z, flag, it
    tz = torch.from_numpy(np.ascontiguousarray(z)).to(device)
    tf = torch.from_numpy(np.ascontiguousarray(flag.astype(np.int32))).to(device)
    ti = torch.from_numpy(np.ascontiguousarray(it.astype(np.int32))).to(device)
    gz, gf, gi = all_gather_results(tz, tf, ti, batch.B, group)
    return gz.cpu().numpy(), gf.cpu().numpy(), gi.cpu().numpy()
