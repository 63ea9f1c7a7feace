"""GPU tests of the optimizer-level pieces added in round 2: the whole L-BFGS solve as one CUDA graph
(LBFGSOpt.optimize_graphed == the eager loop, bit for bit), the seed-sharded solve wrapper (ShardedSolver, world size 1 in
this process; world size 2 over NCCL when two GPUs are visible) and the host-buffer pipeline (HostRolloutPipeline)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import random_q  # noqa: E402
from oracle import rollout_oracle as O  # noqa: E402

DEV = "cuda:0"


def _ik_problem(problems=16, seeds=8, n=4, dev=DEV, lo=0, hi=None):
    from curobo_b200.robot_model import load_robot
    from curobo_b200.rollout import RolloutConfig, RolloutEngine
    from curobo_b200.scene import CuboidData
    from curobo_b200.world import make_benchmark_cuboid_world
    rm = load_robot("franka")
    hi = problems * seeds if hi is None else hi
    _, _, gp, gq = O.fk_forward(rm, random_q(rm, problems, seed=21) * 0.8)
    idx = np.repeat(np.arange(problems * seeds) // seeds, n).astype(np.int32)[lo * n:hi * n]
    eng = RolloutEngine(rm, RolloutConfig.ik(), dev, CuboidData.from_world(make_benchmark_cuboid_world(), dev))
    T = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    eng.update_goal(T(gp[:, :, None, :]), T(gq[:, :, None, :]), T(idx))
    x0 = T(random_q(rm, problems * seeds, seed=22)[lo:hi])
    return rm, eng, x0


def test_graphed_solve_equals_eager_loop():
    from curobo_b200.optim import LBFGSOpt, LBFGSOptCfg
    rm, eng, x0 = _ik_problem()
    B, D, n = x0.shape[0], rm.num_dof, 4

    def cost_grad(x):
        out = eng.evaluate_action(x.view(B * n, 1, D))
        return out.cost.view(-1), out.grad_q.view(B * n, D)

    T = lambda a: torch.as_tensor(a).to(DEV)  # noqa: E731
    opt = LBFGSOpt(LBFGSOptCfg(num_iters=25), B, 1, D, T(rm.position_limits[0]), T(rm.position_limits[1]), cost_grad, DEV)
    q_eager = opt.optimize(x0).clone()
    c_eager = opt.best_cost.clone()
    q_graph = opt.optimize_graphed(x0).clone()          # first call: warm-up + capture + replay
    assert torch.equal(q_graph, q_eager) and torch.equal(opt.best_cost, c_eager)
    x1 = x0.flip(0).contiguous()                        # second call: replay only, on new seeds
    q1 = opt.optimize_graphed(x1).clone()
    assert torch.equal(q1, opt.optimize(x1))
    assert float(opt.best_cost.min()) < float(c_eager.max())


def test_sharded_solver_single_process_matches_plain_solve():
    from curobo_b200.optim import LBFGSOptCfg
    from curobo_b200.sharded import ShardedSolver
    rm, eng, x0 = _ik_problem()
    total = x0.shape[0]
    solver = ShardedSolver(eng, total, 1, LBFGSOptCfg(num_iters=15))
    cost_all, row, best = solver.solve(x0.view(total, 1, -1))
    assert cost_all.shape == (total,) and 0 <= row < total
    assert row == int(torch.argmin(cost_all))
    assert torch.equal(best.view(-1), solver.opt.best_action[row])
    c2, r2, b2 = solver.solve(x0.view(total, 1, -1), graphed=False)
    assert torch.equal(c2, cost_all) and r2 == row and torch.equal(b2, best)


def test_host_pipeline_matches_direct_call():
    from curobo_b200.rollout import HostRolloutPipeline
    rm, eng_a, x0 = _ik_problem(n=1)
    _, eng_b, _ = _ik_problem(n=1)
    _, eng_c, _ = _ik_problem(n=1)
    B = x0.shape[0]
    pipe = HostRolloutPipeline([eng_a, eng_b], B, 1)
    q = [x0.view(B, 1, -1), x0.flip(0).contiguous().view(B, 1, -1)]
    for k in range(2):
        pipe.slots[k].q_host.copy_(q[k].cpu())
        pipe.submit(k)
    for k in range(2):
        cost, grad = pipe.result(k)
        want = eng_c.evaluate_action(q[k])
        assert torch.equal(cost, want.cost.cpu()) and torch.equal(grad, want.grad_q.cpu())
    assert pipe.h2d_bytes == B * rm.num_dof * 4 and pipe.d2h_bytes == B * 4 + B * rm.num_dof * 4


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    from curobo_b200.optim import LBFGSOptCfg
    from curobo_b200.sharded import ShardedSolver, shard_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        problems, seeds = 16, 8
        total = problems * seeds
        lo, hi = shard_rows(total, rank, world)
        _, eng, x0 = _ik_problem(problems, seeds, dev=f"cuda:{rank}", lo=lo, hi=hi)
        solver = ShardedSolver(eng, total, 1, LBFGSOptCfg(num_iters=15))
        cost_all, row, best = solver.solve(x0.view(hi - lo, 1, -1))
        q.put((rank, cost_all.cpu().numpy(), row, best.cpu().numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_solver_two_ranks_equal_one_rank():
    import torch.multiprocessing as mp
    from curobo_b200.optim import LBFGSOptCfg
    from curobo_b200.sharded import ShardedSolver
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    qq = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, qq)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([qq.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    rm, eng, x0 = _ik_problem()
    total = x0.shape[0]
    cost_all, row, best = ShardedSolver(eng, total, 1, LBFGSOptCfg(num_iters=15)).solve(x0.view(total, 1, -1))
    for _, c, r, b in res:                                  # every rank holds the same answer = the single-GPU answer
        np.testing.assert_array_equal(c, cost_all.cpu().numpy())
        assert r == row
        np.testing.assert_array_equal(b, best.cpu().numpy())
