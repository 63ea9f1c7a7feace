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
    if not (dist.is_available() and dist.is_initialized()):       # single process: the gather is the identity
        if cost_local.shape[0] != total_rows:
            raise ValueError(f"single process owns all {total_rows} rows but got {cost_local.shape[0]} costs")
        li = int(torch.argmin(cost_local))
        return cost_local.clone(), li, action_local[li].contiguous()
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


def _rank_world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def local_rows_of(t: torch.Tensor, total_rows: int, group=None) -> torch.Tensor:
    """This rank's slice of a per-seed tensor [total_rows, ...] (GoalRegistry rows: idxs_goal, env_query_idx, seeds;
    rollout/goal_registry.py:27-58)."""
    rank, world = _rank_world(group)
    lo, hi = shard_rows(total_rows, rank, world)
    return t[lo:hi].contiguous()


class ShardedSolver:
    """A seed-sharded solve as ONE call ("ShardedRollout / solver wrapper that slices GoalRegistry rows").

    `total_rows` seeds (particles) of horizon H are split over the ranks of `group`; rank r runs the complete L-BFGS loop
    (curobo_b200.optim.LBFGSOpt: step direction -> fused rollout on rows x line-search candidates -> Wolfe line search)
    on its rows [lo, hi) only -- captured in one CUDA graph -- and the single exchange of the path happens at the end:
    `gather_seed_costs_and_best` (two KB-scale all_gathers over NCCL / NVLink).  No per-iteration communication.

    `engine` must already hold this rank's per-row inputs, repeated per line-search candidate (row order = seed-major,
    candidate-minor, the layout LBFGSOpt evaluates): use `candidate_rows(local_rows_of(idxs_goal, total), n)`.
    `eval_kwargs` are passed to RolloutEngine.evaluate_action on every call (dt, env_query_idx, ...)."""

    def __init__(self, engine, total_rows: int, horizon: int, opt_cfg=None, group=None, eval_kwargs=None):
        from .optim import LBFGSOpt, LBFGSOptCfg
        self.engine, self.total_rows, self.H, self.group = engine, int(total_rows), int(horizon), group
        self.rank, self.world = _rank_world(group)
        self.lo, self.hi = shard_rows(self.total_rows, self.rank, self.world)
        self.b = self.hi - self.lo
        if self.b <= 0:
            raise ValueError(f"rank {self.rank} of {self.world} owns no rows of {total_rows}")
        cfg = opt_cfg if opt_cfg is not None else LBFGSOptCfg()
        rm, dev = engine.robot, engine.device
        self.D = rm.num_dof
        self.n = len(cfg.line_search_scale)
        self._kw = dict(eval_kwargs or {})
        lim = torch.as_tensor(rm.position_limits, dtype=torch.float32, device=dev)
        self.opt = LBFGSOpt(cfg, self.b, self.H, self.D, lim[0].contiguous(), lim[1].contiguous(), self._cost_grad, dev)

    def _cost_grad(self, x: torch.Tensor):
        rows = self.b * self.n
        out = self.engine.evaluate_action(x.view(rows, self.H, self.D), **self._kw)
        cost = out.cost.view(rows) if self.H == 1 else out.cost.sum(dim=1)
        return cost, out.grad_q.view(rows, self.H * self.D)

    def solve(self, x0_local: torch.Tensor, num_iters: Optional[int] = None, graphed: bool = True):
        """x0_local [hi - lo, H, D].  Returns (cost_all [total_rows], best_global_row, best_action [H, D]) -- identical on
        every rank."""
        if x0_local.shape[0] != self.b:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo}, {self.hi}) but got {x0_local.shape[0]} seeds")
        run = self.opt.optimize_graphed if graphed else self.opt.optimize
        q = run(x0_local, num_iters)
        cost_all, row, best = gather_seed_costs_and_best(self.opt.best_cost, q.reshape(self.b, -1), self.total_rows, self.group)
        return cost_all, row, best.view(self.H, self.D)


def candidate_rows(t: torch.Tensor, n_candidates: int) -> torch.Tensor:
    """Repeat a per-seed tensor for the n line-search candidates of each seed (seed-major, candidate-minor)."""
    return t.repeat_interleave(n_candidates, dim=0).contiguous()
