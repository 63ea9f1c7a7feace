"""`B200RobotRollout` (curobo_b200/rollout_protocol.py): the fused rollout behind the reference's `Rollout` Protocol
(rollout/rollout_protocol.py:35-176), consumed the way the reference's optimizer consumes a rollout
(optim/components/gradient_opt_core.py:445-480)."""
import numpy as np
import pytest
import torch

from helpers import random_q, random_walk_q, small_voxel_world
from curobo_b200.robot_model import load_robot
from curobo_b200.rollout import RolloutConfig, RolloutEngine
from curobo_b200.rollout_protocol import B200RobotRollout, RolloutMetrics, RolloutResult
from curobo_b200.scene import CuboidData, VoxelData
from curobo_b200.trajectory import JointState
from curobo_b200.world import make_benchmark_cuboid_world
from oracle import rollout_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# members of the reference Protocol (rollout_protocol.py:46-176)
PROTOCOL_PROPERTIES = ("action_dim", "action_horizon", "action_bound_lows", "action_bound_highs", "dt", "sum_horizon")
PROTOCOL_METHODS = ("evaluate_action", "compute_metrics_from_state", "compute_metrics_from_action", "update_params",
                    "update_batch_size", "update_dt", "reset", "reset_shape", "reset_seed")


def T(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t.to(dt) if dt is not None else t


def optimizer_cost_and_gradient(rollout, x):
    """What GradientOptCore._compute_cost_constraint_and_gradient does with a rollout (gradient_opt_core.py:445-480)."""
    x_n = x.detach().requires_grad_(True)
    res = rollout.evaluate_action(x_n, use_cuda_graph=False)
    costs = res.costs_and_constraints.get_sum_cost_and_constraint(sum_horizon=True)
    assert costs.shape == (x.shape[0],)
    costs.backward(gradient=torch.ones_like(costs), retain_graph=False)
    return costs.detach(), x_n.grad.detach(), res


def _ik_rollout(rm, B, G=8):
    cub = make_benchmark_cuboid_world()
    ro = B200RobotRollout(rm, RolloutConfig.ik(), DEV, cuboid=CuboidData.from_world(cub, DEV), horizon=1)
    _, _, gp, gq = O.fk_forward(rm, random_q(rm, G, seed=5))
    goal = (gp[:, :, None, :].copy(), gq[:, :, None, :].copy())
    idx = (np.arange(B) % G).astype(np.int32)
    ro.update_params(goal_position=T(goal[0]), goal_quat=T(goal[1]), idxs_goal=T(idx))
    return ro, cub, goal, idx


def test_protocol_surface():
    rm = load_robot("franka")
    ro, *_ = _ik_rollout(rm, 8)
    for name in PROTOCOL_PROPERTIES:
        assert isinstance(getattr(type(ro), name), property), name
    for name in PROTOCOL_METHODS:
        assert callable(getattr(ro, name)), name
    assert ro.action_dim == 7 and ro.action_horizon == 1 and ro.sum_horizon
    np.testing.assert_array_equal(ro.action_bound_lows.cpu().numpy(), rm.position_limits[0])
    np.testing.assert_array_equal(ro.action_bound_highs.cpu().numpy(), rm.position_limits[1])


def test_ik_rollout_through_optimizer_contract_vs_oracle():
    rm = load_robot("franka")
    B = 48
    ro, cub, goal, idx = _ik_rollout(rm, B)
    q = random_q(rm, B, seed=6)[:, None, :]
    cost, grad, res = optimizer_cost_and_gradient(ro, T(q))
    assert isinstance(res, RolloutResult) and res.costs_and_constraints.costs.names == ["tool_pose", "cspace"]
    assert res.costs_and_constraints.constraints.names == ["scene_collision", "self_collision"]
    want = O.rollout_cost_grad(rm, q, RolloutConfig.ik().to_oracle_cfg(1), world_cuboid=cub, goal_pos=goal[0], goal_quat=goal[1],
                               idxs_goal=idx)
    np.testing.assert_allclose(cost.cpu().numpy(), want["cost"], rtol=2e-4, atol=1e-5 * want["cost"].max())
    g = want["grad_q"]
    np.testing.assert_allclose(grad.cpu().numpy(), g, rtol=2e-3, atol=2e-5 * np.abs(g).max())
    # the same numbers as the engine called directly
    eng = RolloutEngine(rm, RolloutConfig.ik(), DEV, CuboidData.from_world(cub, DEV))
    eng.update_goal(T(goal[0]), T(goal[1]), T(idx))
    o = eng.evaluate_action(T(q))
    assert torch.equal(o.grad_q, grad) and torch.allclose(o.cost.sum(1), cost, rtol=1e-6)
    # metrics: feasibility = constraints <= 0; convergence = pose errors
    m = ro.compute_metrics_from_action(T(q))
    assert isinstance(m, RolloutMetrics)
    feas_want = (want["scene_cost"].sum(-1) + want["self_cost"]).sum(-1) <= 0
    np.testing.assert_array_equal(m.feasible.cpu().numpy(), feas_want)
    assert m.convergence.names == ["position_tolerance", "orientation_tolerance"]
    np.testing.assert_allclose(m.convergence.values[0].cpu().numpy(), want["pose_pos_err"].reshape(B, 1, -1), rtol=1e-3, atol=1e-5)
    np.testing.assert_allclose(m.convergence.values[1].cpu().numpy(), want["pose_rot_err"].reshape(B, 1, -1), rtol=1e-3, atol=1e-4)
    # batch size change re-allocates once; a second call with the same size does not
    q2 = random_q(rm, 16, seed=7)[:, None, :]
    ro.update_params(idxs_goal=None)
    ro.engine.update_goal(T(goal[0]), T(goal[1]), T(idx[:16].copy()))
    c2, g2, _ = optimizer_cost_and_gradient(ro, T(q2))
    buf = ro.engine.out.cost.data_ptr()
    optimizer_cost_and_gradient(ro, T(q2))
    assert ro.batch_size == 16 and ro.engine.out.cost.data_ptr() == buf and torch.isfinite(g2).all()


def test_mpc_position_actions_with_state():
    """position action space, H = 12, the shipped MPC cost; vel / acc / jerk supplied as the rollout state."""
    rm = load_robot("franka")
    B, H = 6, 12
    cfg = RolloutConfig.mpc()
    cub, vox = make_benchmark_cuboid_world(), small_voxel_world()
    ro = B200RobotRollout(rm, cfg, DEV, cuboid=CuboidData.from_world(cub, DEV), voxel=VoxelData.from_world(vox, DEV), horizon=H)
    q = random_walk_q(rm, B, H, seed=8)
    rng = np.random.default_rng(8)
    v = (np.gradient(q, axis=1) / 0.05).astype(np.float32)
    a_, j_ = rng.normal(0, 5, q.shape).astype(np.float32), rng.normal(0, 200, q.shape).astype(np.float32)
    dt = np.full(B, 0.05, np.float32)
    _, _, gp, gq = O.fk_forward(rm, random_q(rm, B, seed=9))
    goal = (gp[:, :, None, :].copy(), gq[:, :, None, :].copy())
    idx = np.arange(B, dtype=np.int32)
    target = random_q(rm, 1, seed=10, scale=0.4)
    ro.update_params(goal_position=T(goal[0]), goal_quat=T(goal[1]), idxs_goal=T(idx), cspace_target=T(target),
                     state=JointState(T(q), T(v), T(a_), T(j_), T(dt)))
    cost, grad, _ = optimizer_cost_and_gradient(ro, T(q))
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), world_cuboid=cub, world_voxel=vox, goal_pos=goal[0], goal_quat=goal[1],
                               idxs_goal=idx, vel=v, acc=a_, jerk=j_, dt=dt, cspace_target=target)
    np.testing.assert_allclose(cost.cpu().numpy(), want["cost"], rtol=2e-4, atol=1e-5 * want["cost"].max())
    g = want["grad_q"]
    np.testing.assert_allclose(grad.cpu().numpy(), g, rtol=2e-3, atol=2e-5 * np.abs(g).max())


def test_bspline_action_space_gradient_wrt_knots():
    """bspline action space: act_seq = knots; the optimizer contract returns d cost / d knots (fused spline front end + adjoint);
    checked against the engine's evaluate_knots and by a directional finite difference."""
    from bspline_cases import make_case
    rm = load_robot("franka")
    c = make_case(seed=3, B=5, nk=8, D=7, steps=4, degree=4, implicit=False)
    # small weights: the summed cost stays O(1), where a float32 finite difference resolves the gradient; no pose term here --
    # the reference's rotation gradient is hand-defined, not the derivative of its cost (wp_tool_pose.py:113-126)
    cfg = RolloutConfig(pose_weight=None, cspace_type="state", cspace_weight=(1.0, 1e-2, 1e-4, 1e-7, 0.0),
                        cspace_activation=(0.01,) * 5, cspace_reg=(1e-2, 1e-4, 1e-8, 0.0, 0.0), scene_weight=2.0,
                        scene_activation=0.02, use_sweep=True, use_speed_metric=False, self_weight=1.0)
    cub = make_benchmark_cuboid_world()
    ro = B200RobotRollout(rm, cfg, DEV, cuboid=CuboidData.from_world(cub, DEV), action_space="bspline", n_knots=8,
                          bspline_degree=4, interpolation_steps=4)
    assert ro.action_horizon == 8 and ro.horizon == c["T"]
    lo, hi = rm.position_limits
    knots = np.clip(c["knots"] * 0.3, lo + 0.1, hi - 0.1).astype(np.float32)
    mk = lambda arrs, dtv=None: JointState(*[T(np.clip(x * 0.1, lo + 0.1, hi - 0.1) if i == 0 else x * 0.0) for i, x in enumerate(arrs)], dtv)  # noqa: E731
    start, goal_state = mk(c["start"]), mk(c["goal"], T(c["traj_dt"]))
    _, _, gp, gq = O.fk_forward(rm, random_q(rm, 5, seed=11))
    ro.update_params(goal_position=T(gp[:, :, None, :].copy()), goal_quat=T(gq[:, :, None, :].copy()),
                     idxs_goal=T(np.arange(5, dtype=np.int32)), start_state=start, goal_state=goal_state,
                     start_state_idx=T(c["start_idx"]), goal_state_idx=T(c["goal_idx"]), use_implicit_goal_state=T(c["implicit"]))
    x = T(knots)
    cost, grad, res = optimizer_cost_and_gradient(ro, x)
    assert grad.shape == x.shape and torch.isfinite(grad).all() and float(grad.abs().max()) > 0
    assert res.state.position.shape == (5, c["T"], 7)
    # directional derivative of the summed cost along a random direction
    torch.manual_seed(0)
    d = torch.randn_like(x)
    d = d / d.norm()
    eps = 1e-3
    cp, _, _ = optimizer_cost_and_gradient(ro, (x + eps * d).contiguous())
    cm, _, _ = optimizer_cost_and_gradient(ro, (x - eps * d).contiguous())
    fd = float((cp.double().sum() - cm.double().sum()) / (2 * eps))
    an = float((grad.double() * d.double()).sum())
    assert abs(fd - an) <= 0.05 * max(abs(an), abs(fd)) + 2e-6 * float(cost.sum()) / eps, (fd, an, float(cost.sum()))
    # and identical to the engine's own knots entry point
    s_ = ro._spline_args
    o = ro.engine.evaluate_knots(x, s_["start"], s_["start_idx"], s_["goal"], s_["goal_idx"], s_["implicit"], 4, 4)
    assert torch.equal(o.grad_knots, grad)
