"""World-size-2 `gloo` test (CPU) of the N>1 host logic: row sharding + the end-of-solve all_gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from curobo_b200.sharded import candidate_rows, gather_seed_costs_and_best, local_rows_of, shard_rows


def test_shard_rows_partition():
    for total in (0, 1, 7, 16384, 8192 + 3):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(total, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        cost = torch.rand(total, generator=g)
        act = torch.arange(total * 3, dtype=torch.float32).view(total, 3)
        s, e = shard_rows(total, rank, world)
        cost_all, row, best = gather_seed_costs_and_best(cost[s:e].clone(), act[s:e].clone(), total)
        ok = torch.equal(cost_all, cost) and row == int(torch.argmin(cost)) and torch.equal(best, act[row])
        # the GoalRegistry-row slicing the sharded solver uses: this rank's rows, repeated per line-search candidate
        mine = local_rows_of(act, total)
        ok = ok and torch.equal(mine, act[s:e]) and torch.equal(candidate_rows(mine, 4)[::4], mine) \
            and candidate_rows(mine, 4).shape[0] == 4 * (e - s)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [9, 64])
def test_gather_best_world_size_2(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_gather_single_process_is_identity():
    cost = torch.tensor([3.0, 1.0, 2.0])
    act = torch.arange(6, dtype=torch.float32).view(3, 2)
    cost_all, row, best = gather_seed_costs_and_best(cost, act, 3)
    assert torch.equal(cost_all, cost) and row == 1 and torch.equal(best, act[1])
    with pytest.raises(ValueError):
        gather_seed_costs_and_best(cost, act, 4)
