"""Seed-sharded multi-GPU rollout (new functionality; the reference is single-GPU).

Every (seed x waypoint) evaluation is independent, and the only coupling is temporal within one seed's
trajectory, so the batch is sharded on B (never on H): one process per GPU, rank r owns the seed rows
[r*B/W, (r+1)*B/W); robot constants and the world (KBs + the 32 MiB ESDF) are replicated.  There is NO
data-path collective per optimizer iteration.  The only exchange is at the end of a solve: one
all_gather of the per-seed costs plus each rank's best action (KB-scale, latency bound) over
NCCL/NVLink (gloo on CPU for the host-logic tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(total_rows: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of `rank` (first `total_rows % world_size` ranks get one extra)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(total_rows, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_seed_costs_and_best(cost_local: torch.Tensor, action_local: torch.Tensor, total_rows: int,
                               group: Optional[dist.ProcessGroup] = None):
    """End-of-solve exchange.

    cost_local [b_local] per-seed final cost, action_local [b_local, ...] per-seed action of this rank.
    Returns (cost_all [total_rows], best_global_row, best_action) identical on every rank.
    One all_gather of padded costs + one all_gather of each rank's local-best action."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    s, e = shard_rows(total_rows, rank, world)
    if cost_local.shape[0] != e - s:
        raise ValueError(f"rank {rank} owns {e - s} rows but got {cost_local.shape[0]} costs")
    width = (total_rows + world - 1) // world
    pad = torch.full((width,), float("inf"), dtype=cost_local.dtype, device=cost_local.device)
    pad[: e - s] = cost_local
    gathered = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(gathered, pad, group=group)
    cost_all = torch.cat([g[: shard_rows(total_rows, r, world)[1] - shard_rows(total_rows, r, world)[0]]
                          for r, g in enumerate(gathered)])
    if e - s > 0:
        li = int(torch.argmin(cost_local))
        best_local = action_local[li].contiguous()
    else:
        best_local = torch.zeros(action_local.shape[1:], dtype=action_local.dtype, device=action_local.device)
    bests = [torch.empty_like(best_local) for _ in range(world)]
    dist.all_gather(bests, best_local, group=group)
    best_row = int(torch.argmin(cost_all))
    owner = next(r for r in range(world) if shard_rows(total_rows, r, world)[0] <= best_row < shard_rows(total_rows, r, world)[1])
    return cost_all, best_row, bests[owner]
