"""GPU parity of the dynamics-aware STATE cost (second half): RNEA -> effort channel of the c-space
STATE kernel -> RNEA adjoint, composed on the host by curobo_b200.dynamics.DynamicsStateCost, against the composition of the
two oracles (pinned on the CPU by finite differences, tests/test_dynamics_cpu.py).  Runs last in the suite on purpose: it
is the newest composition."""
import numpy as np
import pytest
import torch

from dynamics_cases import effort_cost_oracle, effort_cost_setup
from curobo_b200.dynamics import Dynamics, DynamicsStateCost

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("robot,B,H", [("franka", 6, 5), ("g1_29", 3, 4)])
def test_dynamics_state_cost_vs_oracle(robot, B, H):
    c, shape, jerk, dt, limits, weight, act, reg = effort_cost_setup(robot, B, H)
    want_c, want_g, want_tau = effort_cost_oracle(c, shape, jerk, dt, limits, weight, act, reg)
    dyn = Dynamics(c["rm"], c["mc"], c["inn"], gravity=(0.0, 0.0, -9.81), device=DEV)
    op = DynamicsStateCost(dyn, limits, weight, act, reg)
    r = lambda x: T(np.asarray(x, np.float32).reshape(shape))  # noqa: E731
    q, qd, qdd = r(c["q"]), r(c["qd"]), r(c["qdd"])
    cost, gp, gv, ga, gj, tau = op.evaluate(q, qd, qdd, T(jerk), T(dt))
    torch.cuda.synchronize()

    def close(got, want, rtol):
        got = got.cpu().numpy()
        assert np.allclose(got, want, rtol=rtol, atol=rtol * max(float(np.abs(want).max()), 1e-30)), float(np.abs(got - want).max())

    close(tau, want_tau, 1e-4)
    close(cost, want_c, 5e-4)
    for got, want in zip((gp, gv, ga, gj), want_g):
        close(got, want, 2e-3)
    # the effort channel is live in this case (hinge + L2 + energy), and a second call reuses the buffers bit for bit
    assert float(np.abs(want_g[2]).max()) > 0
    first = [x.clone() for x in (cost, gp, gv, ga, gj)]
    again = op.evaluate(q, qd, qdd, T(jerk), T(dt))
    torch.cuda.synchronize()
    for a_, b_ in zip(first, again[:5]):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("robot,B,H", [("franka", 3, 7), ("g1_29", 2, 4)])
def test_dynamics_aware_rollout_vs_oracle(robot, B, H):
    """RolloutEngine.attach_dynamics: the fused trajectory rollout (swept ESDF + cuboid collision, speed metric, pose, STATE c-space)
    plus the effort channel fed by RNEA.  Oracle = rollout oracle (torque-free) + the effort-only STATE cost composed with the RNEA
    oracle and its adjoint; costs and the gradients w.r.t. position, velocity and acceleration must be the sums."""
    from curobo_b200.rollout import RolloutConfig, RolloutEngine
    from curobo_b200.scene import CuboidData, VoxelData
    from curobo_b200.world import make_benchmark_cuboid_world
    from dynamics_cases import make_case, model_args
    from helpers import random_q, random_walk_q, small_voxel_world
    from oracle import dynamics_oracle as do
    from oracle import rollout_oracle as O
    c = make_case(robot, B * H, 31)
    rm = c["rm"]
    D = c["D"]
    q = random_walk_q(rm, B, H, seed=71)
    rng = np.random.default_rng(3)
    dt = np.full(B, 0.05, np.float32)
    v = (np.gradient(q, axis=1) / 0.05).astype(np.float32)
    a_ = rng.normal(0, 3.0, size=q.shape).astype(np.float32)
    j_ = rng.normal(0, 100.0, size=q.shape).astype(np.float32)
    m = model_args(c)
    tau, cache = do.rnea_forward(q.reshape(-1, D), v.reshape(-1, D), a_.reshape(-1, D), *m)
    elim = np.stack([np.quantile(tau, 0.25, axis=0), np.quantile(tau, 0.75, axis=0)]).astype(np.float32)
    cfg = RolloutConfig.trajopt()
    cfg.cspace_reg = (1000.0, 10000.0, 5.0, 0.05, 40.0)
    cfg.cspace_activation = (0.01, 0.01, 0.01, 0.01, 0.05)
    cub, vox = make_benchmark_cuboid_world(), small_voxel_world()
    _, _, p, qt = O.fk_forward(rm, random_q(rm, B, seed=72))
    gp, gq = p[:, :, None, :].copy(), qt[:, :, None, :].copy()
    idx = np.arange(B, dtype=np.int32)
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    eng.update_goal(T(gp), T(gq), T(idx), non_terminal_axes=torch.zeros((rm.num_tool_frames, 6), dtype=torch.float32, device=DEV))
    base = eng.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))
    base_cost, base_gq = base.cost.clone(), base.grad_q.clone()
    eng.attach_dynamics(Dynamics(rm, c["mc"], c["inn"], gravity=(0.0, 0.0, -9.81), device=DEV), effort_limits=elim)
    out = eng.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))
    torch.cuda.synchronize()
    Lt = rm.num_tool_frames
    ocfg = cfg.to_oracle_cfg(Lt)
    ocfg["pose_non_terminal_axes"] = np.zeros((Lt, 6), np.float32)
    want = O.rollout_cost_grad(rm, q, ocfg, world_cuboid=cub, world_voxel=vox, goal_pos=gp, goal_quat=gq, idxs_goal=idx, vel=v,
                               acc=a_, jerk=j_, dt=dt)
    lim = dict(p=rm.position_limits, v=rm.velocity_limits, a=rm.acceleration_limits, j=rm.jerk_limits, tau=elim)
    w_eff = np.array([0, 0, 0, 0, cfg.cspace_weight[4]], np.float32)
    r_eff = np.array([0, 0, 0, cfg.cspace_reg[3], cfg.cspace_reg[4]], np.float32)
    ec, eg = O.cspace_state_cost(q, v, a_, j_, dt, lim, w_eff, np.asarray(cfg.cspace_activation, np.float32), r_eff, True, True,
                                 effort=tau.reshape(B, H, D))
    bq, bqd, bqdd = do.rnea_backward(eg[4].reshape(-1, D), q.reshape(-1, D), v.reshape(-1, D), cache, *m)
    r = lambda x: np.asarray(x, np.float32).reshape(B, H, D)  # noqa: E731
    assert float(ec.sum()) > 1e-3 * float(want["cost_bh"].sum()), "the effort terms must matter in this case"
    total = want["cost_bh"] + ec.sum(-1)
    assert np.allclose(out.cost.cpu().numpy(), total, rtol=3e-4, atol=1e-5 * total.max())
    assert not torch.equal(out.cost, base_cost) and not torch.equal(out.grad_q, base_gq)
    g = want["grad_q"] + eg[0] + r(bq)
    assert np.allclose(out.grad_q.cpu().numpy(), g, rtol=3e-3, atol=3e-5 * np.abs(g).max())
    gv = want["cspace_grads"][1] + eg[1] + r(bqd)
    ga = want["cspace_grads"][2] + eg[2] + r(bqdd)
    assert np.allclose(out.grad_vel.cpu().numpy(), gv, rtol=3e-3, atol=3e-5 * np.abs(gv).max())
    assert np.allclose(out.grad_acc.cpu().numpy(), ga, rtol=3e-3, atol=3e-5 * np.abs(ga).max())
    # the same terms evaluated INSIDE the trajectory kernel (cb200_rollout_io.dynamics): the effort limits are the blob's, so the
    # engine is built on a robot model that carries the limits of this case
    import dataclasses
    rm2 = dataclasses.replace(rm, effort_limits=elim)
    eng2 = RolloutEngine(rm2, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    eng2.update_goal(T(gp), T(gq), T(idx), non_terminal_axes=torch.zeros((rm.num_tool_frames, 6), dtype=torch.float32, device=DEV))
    eng2.attach_dynamics(Dynamics(rm2, c["mc"], c["inn"], gravity=(0.0, 0.0, -9.81), device=DEV), fused=True)
    assert eng2._dyn_params is not None and eng2._effort_cost is None
    fused = eng2.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))
    torch.cuda.synchronize()
    assert np.allclose(fused.cost.cpu().numpy(), total, rtol=3e-4, atol=1e-5 * total.max())
    assert np.allclose(fused.grad_q.cpu().numpy(), g, rtol=3e-3, atol=3e-5 * np.abs(g).max())
    assert np.allclose(fused.grad_vel.cpu().numpy(), gv, rtol=3e-3, atol=3e-5 * np.abs(gv).max())
    assert np.allclose(fused.grad_acc.cpu().numpy(), ga, rtol=3e-3, atol=3e-5 * np.abs(ga).max())
    assert np.allclose(fused.cspace_cost.cpu().numpy(), want["cspace_cost"] + ec, rtol=3e-4, atol=1e-5 * float((want["cspace_cost"] + ec).max()))


def test_dynamics_aware_knots_rollout_is_consistent():
    """evaluate_knots with the in-kernel inverse dynamics (expanded spline schedule): cost and d cost / d knots equal the chain
    spline -> dynamics-aware evaluate_action on the spline states -> spline adjoint, i.e. the RNEA-adjoint gradients reach the knots."""
    import dataclasses
    from curobo_b200.backends import trajectory as trajectory_cu
    from curobo_b200.rollout import RolloutConfig, RolloutEngine
    from curobo_b200.scene import CuboidData, VoxelData
    from curobo_b200.trajectory import JointState
    from curobo_b200.world import make_benchmark_cuboid_world
    from dynamics_cases import make_case
    from helpers import random_q, random_walk_q, small_voxel_world
    from oracle import rollout_oracle as O
    B, nk, degree, steps = 2, 8, 4, 2
    H = (nk + degree + 1) * steps + 1
    c = make_case("franka", 4, 33)
    rm = dataclasses.replace(c["rm"], effort_limits=np.stack([np.full(c["D"], -8.0), np.full(c["D"], 8.0)]).astype(np.float32))
    D = c["D"]
    knots = random_walk_q(rm, B, nk, seed=81)
    z = np.zeros((B, D), np.float32)
    cfg = RolloutConfig.trajopt()
    cfg.cspace_reg = (1000.0, 10000.0, 5.0, 0.05, 40.0)
    cub, vox = make_benchmark_cuboid_world(), small_voxel_world()
    _, _, p, qt = O.fk_forward(rm, random_q(rm, B, seed=82))

    def engine():
        e = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
        e.update_goal(T(p[:, :, None, :].copy()), T(qt[:, :, None, :].copy()), T(np.arange(B, dtype=np.int32)),
                      non_terminal_axes=torch.zeros((1, 6), dtype=torch.float32, device=DEV))
        e.attach_dynamics(Dynamics(rm, c["mc"], c["inn"], gravity=(0.0, 0.0, -9.81), device=DEV), fused=True)
        return e

    ks = JointState(T(knots[:, 0].copy()), T(z), T(z), T(z))
    kg = JointState(T(knots[:, -1].copy()), T(z), T(z), T(z), dt=T(np.full(B, 0.05, np.float32)))
    kidx = torch.arange(B, dtype=torch.int32, device=DEV)
    kimp = torch.zeros(B, dtype=torch.uint8, device=DEV)
    e1 = engine()
    o1 = e1.evaluate_knots(T(knots), ks, kidx, kg, kidx, kimp, degree, steps)
    torch.cuda.synchronize()
    cost1, gk1 = o1.cost.clone(), o1.grad_knots.clone()
    st = [x.clone() for x in e1._state]
    e2 = engine()
    o2 = e2.evaluate_action(st[0], vel=st[1], acc=st[2], jerk=st[3], dt=e1._state_dt.clone())
    gk2 = torch.zeros_like(gk1)
    trajectory_cu.launch_bspline_interpolation_backward_kernel(gk2, o2.grad_q, o2.grad_vel, o2.grad_acc, o2.grad_jerk, kg.dt, kidx,
                                                               kimp, B, H, D, nk, degree)
    torch.cuda.synchronize()
    assert torch.allclose(cost1, o2.cost, rtol=1e-5, atol=1e-6 * float(o2.cost.abs().max()))
    assert torch.allclose(gk1, gk2, rtol=1e-4, atol=1e-6 * float(gk2.abs().max()))
    e3 = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    e3.update_goal(T(p[:, :, None, :].copy()), T(qt[:, :, None, :].copy()), T(np.arange(B, dtype=np.int32)),
                   non_terminal_axes=torch.zeros((1, 6), dtype=torch.float32, device=DEV))
    o3 = e3.evaluate_knots(T(knots), ks, kidx, kg, kidx, kimp, degree, steps)
    assert not torch.allclose(o3.grad_knots, gk1), "the dynamics terms must reach the knots"
    with pytest.raises(ValueError):
        e1.evaluate_knots(T(knots), ks, kidx, kg, kidx, kimp, degree, steps, in_kernel_spline=True)
