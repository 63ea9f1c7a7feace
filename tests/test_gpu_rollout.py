"""GPU parity of the FUSED rollout cost+gradient kernel (through the C ABI) against the oracle and the
golden fixtures at small sizes, and through size-independent properties at BASELINE.json's full sizes.

Tolerances as in test_gpu_parity.py: costs rel 1e-4 (+1e-6*max), grad_q rel 1e-3 + 1e-5*|g|_inf."""
import os

import numpy as np
import pytest
import torch

from helpers import humanoid_q, random_q, random_walk_q, small_voxel_world
from curobo_b200.kinematics import Kinematics, SelfCollisionCost
from curobo_b200.robot_model import load_robot
from curobo_b200.rollout import FusedRolloutFunction, RolloutConfig, RolloutEngine
from curobo_b200.scene import (CollisionBuffer, CuboidData, SceneData, SphereObstacleCollision, VoxelData)
from curobo_b200.world import VoxelWorld, make_benchmark_cuboid_world, make_box_esdf
from oracle import rollout_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def T(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t.to(dt) if dt is not None else t


def grad_close(a, b, rtol=1e-3, scale=1e-5):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=scale * max(np.abs(b).max(), 1e-6))


def cost_close(a, b, rtol=1e-4):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-6 * max(np.abs(b).max(), 1e-6))


def goal_from_q(rm, qg):
    _, _, p, qt = O.fk_forward(rm, qg)
    return p[:, :, None, :].copy(), qt[:, :, None, :].copy()


def last_variant():
    """which kernel the last rollout launch used (include/curobo_b200.h: CB200_VARIANT_*)"""
    from curobo_b200 import lib as cblib
    return int(cblib.load().cb200_last_rollout_variant())


def check_against_oracle(rm, cfg, q, cub=None, vox=None, goal=None, idx=None):
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV) if cub is not None else None,
                        VoxelData.from_world(vox, DEV) if vox is not None else None, store_fk_outputs=True)
    if goal is not None:
        eng.update_goal(T(goal[0]), T(goal[1]), T(idx))
    out = eng.evaluate_action(T(q))
    torch.cuda.synchronize()
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(rm.num_tool_frames), world_cuboid=cub, world_voxel=vox,
                               goal_pos=None if goal is None else goal[0], goal_quat=None if goal is None else goal[1],
                               idxs_goal=idx)
    np.testing.assert_allclose(out.robot_spheres.cpu().numpy(), want["spheres"], atol=1e-5)
    np.testing.assert_allclose(out.link_pos.cpu().numpy(), want["link_pos"], atol=1e-5)
    if "self_cost" in want:
        cost_close(out.self_cost.cpu().numpy(), want["self_cost"])
    if "scene_cost" in want:
        np.testing.assert_allclose(out.scene_cost.cpu().numpy(), want["scene_cost"], rtol=1e-4,
                                   atol=1e-5 * max(want["scene_cost"].max(), 1e-6))
    if "pose_cost" in want:
        np.testing.assert_allclose(out.pose_cost.cpu().numpy(), want["pose_cost"], rtol=2e-4,
                                   atol=1e-5 * want["pose_cost"].max())
    if "cspace_cost" in want:
        cost_close(out.cspace_cost.cpu().numpy(), want["cspace_cost"])
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    return out, want


def test_franka_ik_rollout_vs_oracle_and_golden():
    rm = load_robot("franka")
    g = np.load(os.path.join(GOLD, "franka_ik_rollout_b48.npz"))
    cfg = RolloutConfig.ik()
    out, want = check_against_oracle(rm, cfg, g["q"], cub=make_benchmark_cuboid_world(),
                                     goal=(g["goal_pos"], g["goal_quat"]), idx=g["idxs_goal"])
    np.testing.assert_allclose(out.cost.cpu().numpy(), g["cost_bh"], rtol=2e-4, atol=1e-5 * g["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), g["grad_q"], rtol=2e-3, scale=2e-5)
    assert (want["self_cost"] > 0).any() and (want["scene_cost"] > 0).any()


def test_franka_ik_rollout_larger_batch():
    rm = load_robot("franka")
    q = random_q(rm, 700, seed=41)[:, None, :]
    gp, gq = goal_from_q(rm, random_q(rm, 16, seed=42))
    idx = (np.arange(700) % 16).astype(np.int32)
    check_against_oracle(rm, RolloutConfig.ik(), q, cub=make_benchmark_cuboid_world(), goal=(gp, gq), idx=idx)


def test_franka_esdf_horizon_rollout_terminal_weights():
    """H > 1, discrete ESDF + cuboids; non-terminal pose axes weights zero like the trajopt config."""
    rm = load_robot("franka")
    q = random_walk_q(rm, 10, 6, seed=43)
    cfg = RolloutConfig(self_weight=10000.0, scene_weight=100000.0, scene_activation=0.0025,
                        pose_weight=(1000000.0, 100000.0), cspace_type="position", cspace_weight=(5000.0, 0, 0, 0, 0),
                        cspace_activation=(0.01, 0, 0, 0, 0))
    gp, gq = goal_from_q(rm, random_q(rm, 10, seed=44))
    idx = np.arange(10, dtype=np.int32)
    rmv = small_voxel_world()
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(make_benchmark_cuboid_world(), DEV), VoxelData.from_world(rmv, DEV))
    nt = torch.zeros((1, 6), dtype=torch.float32, device=DEV)
    eng.update_goal(T(gp), T(gq), T(idx), non_terminal_axes=nt)
    out = eng.evaluate_action(T(q))
    ocfg = cfg.to_oracle_cfg(1)
    ocfg["pose_non_terminal_axes"] = np.zeros((1, 6), np.float32)
    want = O.rollout_cost_grad(rm, q, ocfg, world_cuboid=make_benchmark_cuboid_world(), world_voxel=rmv, goal_pos=gp,
                               goal_quat=gq, idxs_goal=idx)
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    assert float(out.pose_cost[:, :-1].abs().sum()) == 0.0 and float(out.pose_cost[:, -1].abs().sum()) > 0


@pytest.mark.parametrize("robot,n", [("g1_29", 6), ("g1_29", 40), ("g1_43", 12)])
def test_humanoid_esdf_rollout_vs_oracle(robot, n):
    rm = load_robot(robot)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0))
    if robot == "g1_29" and n == 6:
        g = np.load(os.path.join(GOLD, "g1_29_esdf_rollout_b6.npz"))
        q = g["q"]
    else:
        q = humanoid_q(rm, n, seed=45)[:, None, :]
    gp, gq = goal_from_q(rm, humanoid_q(rm, 3, seed=46, scale=0.5))
    idx = (np.arange(q.shape[0]) % 3).astype(np.int32)
    cfg_pose = RolloutConfig(**{**cfg.__dict__, "pose_weight": (2000.0, 100.0)})
    check_against_oracle(rm, cfg_pose, q, vox=small_voxel_world(), goal=(gp, gq), idx=idx)
    if robot == "g1_29" and n == 6:
        out, _ = check_against_oracle(rm, cfg, q, vox=small_voxel_world())
        grad_close(out.grad_q.cpu().numpy(), g["grad_q"], rtol=2e-3, scale=2e-5)


def test_state_cspace_rollout_vs_oracle():
    rm = load_robot("franka")
    B, H = 6, 9
    q = random_walk_q(rm, B, H, seed=47)
    rng = np.random.default_rng(1)
    v, a, j = [rng.normal(0, s, size=(B, H, 7)).astype(np.float32) for s in (2.0, 12.0, 400.0)]
    dt = rng.uniform(0.02, 0.1, size=B).astype(np.float32)
    cfg = RolloutConfig(self_weight=10000.0, cspace_type="state", cspace_weight=(10000.0, 10000.0, 100.0, 50.0, 100.0),
                        cspace_activation=(0.01,) * 5, cspace_reg=(1000.0, 10000.0, 5.0, 0.0, 0.0))
    eng = RolloutEngine(rm, cfg, DEV)
    out = eng.evaluate_action(T(q), vel=T(v), acc=T(a), jerk=T(j), dt=T(dt))
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), vel=v, acc=a, jerk=j, dt=dt)
    cost_close(out.cspace_cost.cpu().numpy(), want["cspace_cost"], rtol=2e-4)
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    for got, wantg in zip((out.grad_vel, out.grad_acc, out.grad_jerk), want["cspace_grads"][1:4]):
        grad_close(got.cpu().numpy(), wantg)


# ------------------------------------------------------------------------------------------------
# full-size, size-independent properties (BASELINE.json configs 2 and 5)
# ------------------------------------------------------------------------------------------------
def _ik_engine(rm, B, seed=50):
    gp, gq = goal_from_q(rm, random_q(rm, 512, seed=seed))
    eng = RolloutEngine(rm, RolloutConfig.ik(), DEV, CuboidData.from_world(make_benchmark_cuboid_world(), DEV))
    eng.update_goal(T(gp), T(gq), T((np.arange(B) // 32 % 512).astype(np.int32)))
    return eng


def test_full_size_ik_properties():
    """config 2: 512 targets x 32 seeds.  (i) bitwise run-to-run determinism, (ii) fused == composition of
    the per-op kernels + autograd, (iii) row permutation equivariance, (iv) CUDA-graph replay identical."""
    rm = load_robot("franka")
    B = 512 * 32
    q = T(random_q(rm, B, seed=51)[:, None, :])
    eng = _ik_engine(rm, B)
    o = eng.evaluate_action(q)
    c1, g1 = o.cost.clone(), o.grad_q.clone()
    o = eng.evaluate_action(q)
    assert torch.equal(c1, o.cost) and torch.equal(g1, o.grad_q)
    assert torch.isfinite(c1).all() and torch.isfinite(g1).all()
    # (ii) unfused composition with our per-op kernels (FK -> self + scene -> autograd backward)
    kin = Kinematics(rm, DEV)
    selfc = SelfCollisionCost(rm, eng.cfg.self_weight, DEV)
    scene = SceneData(CuboidData.from_world(make_benchmark_cuboid_world(), DEV), None)
    buf = CollisionBuffer.from_shape((B, 1, rm.num_spheres, 4), DEV)
    qg = q.clone().requires_grad_(True)
    st = kin.compute_kinematics(qg)
    d_self = selfc.forward(st.robot_spheres)
    d_scene = SphereObstacleCollision.apply(st.robot_spheres, buf, scene, T(np.array([eng.cfg.scene_weight], np.float32)),
                                            T(np.array([0.0], np.float32)), None, torch.zeros(B, dtype=torch.int32, device=DEV),
                                            False, False)
    (d_self.sum() + d_scene.sum()).backward()
    eng2 = RolloutEngine(rm, RolloutConfig(self_weight=eng.cfg.self_weight, scene_weight=eng.cfg.scene_weight), DEV,
                         CuboidData.from_world(make_benchmark_cuboid_world(), DEV))
    o2 = eng2.evaluate_action(q)
    torch.testing.assert_close(o2.self_cost.view(-1), d_self.view(-1), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(o2.scene_cost, d_scene, rtol=1e-5, atol=1e-4)
    scale = float(qg.grad.abs().max())
    torch.testing.assert_close(o2.grad_q, qg.grad, rtol=2e-3, atol=2e-5 * scale)
    # (iii) permutation equivariance
    perm = torch.randperm(B, device=DEV)
    idx_perm = eng._goal[2][perm].contiguous()
    eng.update_goal(eng._goal[0], eng._goal[1], idx_perm)
    o3 = eng.evaluate_action(q[perm].contiguous())
    assert torch.equal(o3.cost, c1[perm]) and torch.equal(o3.grad_q, g1[perm])
    # (iv) graph capture + replay
    eng = _ik_engine(rm, B)
    eng.evaluate_action(q)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        eng.evaluate_action(q)
    eng.out.cost.zero_()
    eng.out.grad_q.zero_()
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(eng.out.cost, c1) and torch.equal(eng.out.grad_q, g1)


def test_full_size_humanoid_esdf_properties():
    """config 5 (per-GPU share 1024 seeds of 8192) on the 256^3 fp16 ESDF: determinism, zero weights -> zero
    cost/grad, gradient is a descent direction (cost decreases along -grad for a small step)."""
    rm = load_robot("g1_29")
    sdf = make_box_esdf(n=256, voxel_size=0.01, num_boxes=12, seed=0, xp=torch)
    vox = VoxelData(T(np.array([[[256, 256, 256, 0.01]]], np.float32)), T(np.array([[[0, 0, 0, 1, 0, 0, 0, 0]]], np.float32)),
                    torch.ones((1, 1), dtype=torch.uint8, device=DEV), torch.ones(1, dtype=torch.int32, device=DEV),
                    sdf.reshape(1, 1, -1).contiguous(), 1, 1, 100.0)
    B = 1024
    q = T(humanoid_q(rm, B, seed=52, scale=0.5)[:, None, :])
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0))
    eng = RolloutEngine(rm, cfg, DEV, voxel=vox)
    o = eng.evaluate_action(q)
    c1, g1 = o.cost.clone(), o.grad_q.clone()
    o = eng.evaluate_action(q)
    assert torch.equal(c1, o.cost) and torch.equal(g1, o.grad_q)
    assert torch.isfinite(g1).all() and float(c1.sum()) > 0
    # scene term alone: small step against the gradient does not increase the (piecewise smooth) scene cost
    eng_s = RolloutEngine(rm, RolloutConfig(scene_weight=5000.0, scene_activation=0.02), DEV, voxel=vox)
    o = eng_s.evaluate_action(q)
    cs, gs = o.cost.clone(), o.grad_q.clone()
    hit = cs.view(-1) > 0
    assert int(hit.sum()) > 10
    step = 1e-4 / gs.norm(dim=-1, keepdim=True).clamp(min=1e-9)
    c_new = eng_s.evaluate_action((q - step * gs).contiguous()).cost
    frac = float(((c_new.view(-1) <= cs.view(-1) + 1e-3)[hit]).float().mean())
    assert frac > 0.95, frac
    # zero weights -> exact zeros
    eng_z = RolloutEngine(rm, RolloutConfig(), DEV, voxel=vox)
    o = eng_z.evaluate_action(q)
    assert float(o.cost.abs().sum()) == 0.0 and float(o.grad_q.abs().sum()) == 0.0
    # oracle spot check on a sample of rows at full ESDF size
    sel = np.arange(0, B, 64)
    vw = VoxelWorld(vox.params.cpu().numpy(), vox.inv_pose.cpu().numpy(), np.ones((1, 1), np.uint8), np.ones(1, np.int32),
                    sdf.reshape(1, 1, -1).cpu().numpy(), 100.0)
    want = O.rollout_cost_grad(rm, q.cpu().numpy()[sel], cfg.to_oracle_cfg(14), world_voxel=vw)
    np.testing.assert_allclose(c1.cpu().numpy()[sel], want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(g1.cpu().numpy()[sel], want["grad_q"], rtol=2e-3, scale=2e-5)


def test_autograd_function_and_errors():
    rm = load_robot("franka")
    eng = _ik_engine(rm, 64)
    q = T(random_q(rm, 64, seed=53)[:, None, :]).requires_grad_(True)
    cost = FusedRolloutFunction.apply(q, eng)
    cost.sum().backward()
    assert cost.shape == (64,) and torch.equal(q.grad, eng.out.grad_q)
    with pytest.raises(ValueError):
        eng.evaluate_action(torch.zeros((64, 1, 6), device=DEV))
    with pytest.raises(ValueError):
        eng.evaluate_action(torch.zeros((64, 1, 7)))                       # CPU tensor: no CPU path
    with pytest.raises(ValueError):
        eng.evaluate_action(torch.zeros((64, 2, 7), device=DEV)[:, ::2])   # non-contiguous
    with pytest.raises(ValueError):
        RolloutEngine(rm, RolloutConfig.ik(), "cpu")


# ------------------------------------------------------------------------------------------------
# trajectory mode (swept scene collision + speed metric): rollout_traj_kernel
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H", [(4, 12), (3, 30), (5, 5), (2, 81), (6, 1), (3, 8), (3, 9)])
@pytest.mark.parametrize("speed", [True, False])
def test_traj_rollout_vs_oracle(B, H, speed):
    """configs 3/4 shape: Franka trajectories, 64^3 ESDF + cuboids, swept (3 steps each side) + speed metric,
    STATE c-space with retimed weights, terminal-only pose cost (lbfgs_bspline_trajopt.yml)."""
    rm = load_robot("franka")
    q = random_walk_q(rm, B, H, seed=60 + H)
    rng = np.random.default_rng(H)
    dt = np.full(B, 0.05, np.float32)
    v = np.gradient(q, axis=1).astype(np.float32) / 0.05 if H > 1 else np.zeros_like(q)
    a_ = rng.normal(0, 5.0, size=q.shape).astype(np.float32)
    j_ = rng.normal(0, 200.0, size=q.shape).astype(np.float32)
    cfg = RolloutConfig.trajopt()
    cfg.use_speed_metric = speed
    cub, vox = make_benchmark_cuboid_world(), small_voxel_world()
    gp, gq = goal_from_q(rm, random_q(rm, B, seed=61))
    idx = np.arange(B, dtype=np.int32)
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    nt = torch.zeros((1, 6), dtype=torch.float32, device=DEV)
    eng.update_goal(T(gp), T(gq), T(idx), non_terminal_axes=nt)
    out = eng.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))
    torch.cuda.synchronize()
    ocfg = cfg.to_oracle_cfg(1)
    ocfg["pose_non_terminal_axes"] = np.zeros((1, 6), np.float32)
    want = O.rollout_cost_grad(rm, q, ocfg, world_cuboid=cub, world_voxel=vox, goal_pos=gp, goal_quat=gq, idxs_goal=idx,
                               vel=v, acc=a_, jerk=j_, dt=dt)
    np.testing.assert_allclose(out.scene_cost.cpu().numpy(), want["scene_cost"], rtol=2e-4,
                               atol=1e-5 * max(want["scene_cost"].max(), 1e-6))
    cost_close(out.self_cost.cpu().numpy(), want["self_cost"])
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    # determinism
    c1, g1 = out.cost.clone(), out.grad_q.clone()
    out = eng.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))
    assert torch.equal(c1, out.cost) and torch.equal(g1, out.grad_q)


@pytest.mark.parametrize("B,H", [(40, 20), (700, 12)])
def test_traj_dense_gradients_over_many_tiles(monkeypatch, B, H):
    """Trajectory kernel on a robot BURIED in a solid ball (every sphere collides: the J^T takes the dense, out-of-line form) with
    more tiles than resident CTAs, so CTAs iterate over tiles handed out by the ticket counter: oracle parity on a subset of
    seeds, bit-equality with static striding (CB200_QUEUE=0), determinism, and the ticket counter re-armed."""
    from curobo_b200.world import VoxelWorld
    if DEV == "cpu" and B > 100:
        pytest.skip("full-size variant: GPU only")
    rm = load_robot("franka")
    q = random_walk_q(rm, B, H, seed=80 + H)
    dt = np.full(B, 0.05, np.float32)
    g = np.stack(np.meshgrid(*[np.arange(48)] * 3, indexing="ij"), -1).astype(np.float32)
    vox = VoxelWorld.from_grid((np.linalg.norm(g - 23.5, axis=-1) * 0.06 - 1.2).astype(np.float32), 0.06)
    cfg = RolloutConfig.trajopt()
    gp, gq = goal_from_q(rm, random_q(rm, 2, seed=81))
    idx = (np.arange(B) % 2).astype(np.int32)
    outs = {}
    for name, env in (("queue", {}), ("static", {"CB200_QUEUE": "0"})):
        monkeypatch.delenv("CB200_QUEUE", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng = RolloutEngine(rm, cfg, DEV, None, VoxelData.from_world(vox, DEV))
        eng.update_goal(T(gp), T(gq), T(idx), non_terminal_axes=torch.zeros((1, 6), dtype=torch.float32, device=DEV))
        for _ in range(2):
            out = eng.evaluate_action(T(q), dt=T(dt))
        torch.cuda.synchronize()
        outs[name] = (out.cost.clone(), out.grad_q.clone(), out.scene_cost.clone())
        assert int(eng._work_counter.abs().sum()) == 0
    for x, y in zip(outs["queue"], outs["static"]):
        assert torch.equal(x, y)
    n = min(B, 12)                                                            # oracle on the first seeds
    ocfg = cfg.to_oracle_cfg(1)
    ocfg["pose_non_terminal_axes"] = np.zeros((1, 6), np.float32)
    want = O.rollout_cost_grad(rm, q[:n], ocfg, world_voxel=vox, goal_pos=gp, goal_quat=gq, idxs_goal=idx[:n], dt=dt[:n])
    assert int((want["scene_cost"] > 0).sum(-1).min()) > 2 * rm.num_links, "every row must take the dense J^T"
    cost, grad, _ = outs["queue"]
    np.testing.assert_allclose(cost[:n].cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(grad[:n].cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)


def test_traj_full_size_mpc_properties():
    """config 4 shape: 1024 particles x 30 horizon, swept + speed metric on the 256^3 ESDF.
    fused == FK kernel + swept scene kernel (per-op) on the scene term; stationary trajectories == discrete."""
    from curobo_b200.scene import SweptSphereObstacleCollision
    rm = load_robot("franka")
    B, H = 1024, 30
    sdf = make_box_esdf(n=256, voxel_size=0.01, num_boxes=12, seed=0, xp=torch)
    vox = VoxelData(T(np.array([[[256, 256, 256, 0.01]]], np.float32)), T(np.array([[[0, 0, 0, 1, 0, 0, 0, 0]]], np.float32)),
                    torch.ones((1, 1), dtype=torch.uint8, device=DEV), torch.ones(1, dtype=torch.int32, device=DEV),
                    sdf.reshape(1, 1, -1).contiguous(), 1, 1, 100.0)
    q = T(random_walk_q(rm, B, H, seed=70))
    dt = torch.full((B,), 0.05, dtype=torch.float32, device=DEV)
    cfg = RolloutConfig(scene_weight=100000.0, scene_activation=0.0025, use_sweep=True, use_speed_metric=True)
    eng = RolloutEngine(rm, cfg, DEV, voxel=vox)
    o = eng.evaluate_action(q, dt=dt)
    kin = Kinematics(rm, DEV)
    st = kin.compute_kinematics(q)
    buf = CollisionBuffer.from_shape((B, H, rm.num_spheres, 4), DEV)
    d = SweptSphereObstacleCollision.apply(st.robot_spheres, buf, SceneData(None, vox), T(np.array([100000.0], np.float32)),
                                           T(np.array([0.0025], np.float32)), None, dt[:1].contiguous(), True,
                                           torch.zeros(B, dtype=torch.int32, device=DEV), False, False)
    assert float(d.sum()) > 0
    torch.testing.assert_close(o.scene_cost, d, rtol=1e-4, atol=1e-5 * float(d.max()))
    # stationary trajectories: swept == discrete (reference property test_voxel_collision.py:1014)
    qs = q[:, :1].expand(B, H, rm.num_dof).contiguous()
    o_s = eng.evaluate_action(qs, dt=dt)
    c_sw = o_s.scene_cost.clone()
    eng_d = RolloutEngine(rm, RolloutConfig(scene_weight=100000.0, scene_activation=0.0025), DEV, voxel=vox)
    c_d = eng_d.evaluate_action(qs).scene_cost
    torch.testing.assert_close(c_sw, c_d, rtol=1e-5, atol=1e-6 * float(c_d.max()))


# ------------------------------------------------------------------------------------------------
# options of the fused kernel that the per-op tests cover separately: multi-env worlds, goalsets
# ------------------------------------------------------------------------------------------------
def _two_env_worlds():
    """env 0: benchmark table + pillar and a 64^3 ESDF; env 1: a wall in front of the robot, disabled pillar, and a
    different ESDF."""
    from curobo_b200.world import CuboidWorld
    c0 = make_benchmark_cuboid_world(max_n=4)
    c1 = CuboidWorld.create([{"dims": [0.05, 1.5, 1.5], "pose": [0.35, 0.0, 0.5, 0.9659258, 0, 0, 0.2588190]},
                             {"dims": [0.3, 0.3, 0.3], "pose": [0.0, 0.5, 0.4, 1, 0, 0, 0]},
                             {"dims": [5.0, 5.0, 5.0], "pose": [0.0, 0.0, 0.0, 1, 0, 0, 0]}], max_n=4)
    c1.enable[0, 2] = 0                                    # a disabled obstacle that would hit everything
    cub = CuboidWorld(np.concatenate([c0.dims, c1.dims]), np.concatenate([c0.inv_pose, c1.inv_pose]),
                      np.concatenate([c0.enable, c1.enable]), np.concatenate([c0.count, c1.count]))
    v0, v1 = small_voxel_world(seed=3), small_voxel_world(seed=8, num_boxes=6)
    vox = VoxelWorld(np.concatenate([v0.params, v1.params]), np.concatenate([v0.inv_pose, v1.inv_pose]),
                     np.concatenate([v0.enable, v1.enable]), np.concatenate([v0.count, v1.count]),
                     np.concatenate([v0.features, v1.features]), v0.max_dist)
    return cub, vox


@pytest.mark.parametrize("mode", ["discrete", "swept"])
def test_fused_rollout_multi_env(mode):
    """Rows pick their world through env_query_idx[b] (two envs with different cuboids and ESDFs)."""
    rm = load_robot("franka")
    B, H = 10, (1 if mode == "discrete" else 6)
    q = random_walk_q(rm, B, H, seed=77) if H > 1 else random_q(rm, B, seed=77)[:, None, :]
    cub, vox = _two_env_worlds()
    env = (np.arange(B) % 2).astype(np.int32)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0),
                        use_sweep=(mode == "swept"), use_speed_metric=False)
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    out = eng.evaluate_action(T(q), env_query_idx=T(env))
    torch.cuda.synchronize()
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), world_cuboid=cub, world_voxel=vox, env_query_idx=env)
    assert want["scene_cost"].reshape(B, -1)[0::2].sum() > 0 and want["scene_cost"].reshape(B, -1)[1::2].sum() > 0
    np.testing.assert_allclose(out.scene_cost.cpu().numpy(), want["scene_cost"], rtol=2e-4,
                               atol=1e-5 * max(want["scene_cost"].max(), 1e-6))
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    # the two envs really differ: evaluating every row in env 0 gives another scene cost for the odd rows
    out0 = eng.evaluate_action(T(q), env_query_idx=T(np.zeros(B, np.int32))).scene_cost.clone()
    assert not torch.allclose(out0[1::2], T(want["scene_cost"])[1::2])


@pytest.mark.parametrize("lie", [False, True])
def test_fused_rollout_goalset_and_tool_frames(lie):
    """Goalsets (closest of 3 goals per tool frame) through the fused kernel, on the 14-tool-frame humanoid and the arm."""
    for robot, B in (("franka", 24), ("g1_29", 6)):
        rm = load_robot(robot)
        q = (random_q(rm, B, seed=81) if robot == "franka" else humanoid_q(rm, B, seed=81))[:, None, :]
        G, ngs = 4, 3
        qg = (random_q(rm, G * ngs, seed=82) if robot == "franka" else humanoid_q(rm, G * ngs, seed=82, scale=0.5))
        _, _, p, qt = O.fk_forward(rm, qg)                       # [G*ngs, L, 3/4]
        L = p.shape[1]
        gp = p.reshape(G, ngs, L, 3).transpose(0, 2, 1, 3).copy()
        gq = qt.reshape(G, ngs, L, 4).transpose(0, 2, 1, 3).copy()
        idx = (np.arange(B) % G).astype(np.int32)
        cfg = RolloutConfig(self_weight=0.0, scene_weight=0.0, pose_weight=(2000.0, 100.0), pose_lie=lie,
                            cspace_type="position", cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0))
        eng = RolloutEngine(rm, cfg, DEV, store_fk_outputs=True)
        eng.update_goal(T(gp), T(gq), T(idx))
        out = eng.evaluate_action(T(q))
        torch.cuda.synchronize()
        want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(L), goal_pos=gp, goal_quat=gq, idxs_goal=idx)
        np.testing.assert_allclose(out.pose_cost.cpu().numpy(), want["pose_cost"], rtol=3e-4, atol=2e-5 * want["pose_cost"].max())
        assert np.array_equal(out.pose_goalset_idx.cpu().numpy(), want["pose_goalset_idx"].reshape(B, 1, L))
        assert len(np.unique(want["pose_goalset_idx"])) > 1
        grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=3e-3, scale=3e-5)


# ------------------------------------------------------------------------------------------------
# ESDF lower-bound pyramid level: built on the GPU == numpy construction; every result identical with / without
# ------------------------------------------------------------------------------------------------
def test_voxel_mip_build_and_exact_cull():
    from helpers import mip_block, numpy_voxel_mip
    from curobo_b200.scene import build_voxel_mip
    cub, vox = _two_env_worlds()                       # two 64^3 layers (multiples of 8)
    odd = small_voxel_world(n=45, voxel=0.05, seed=5)  # 45^3: blocks clipped at the upper faces
    for w in (vox, odd):
        vd = VoxelData.from_world(w, DEV)
        mip = build_voxel_mip(vd)
        torch.cuda.synchronize()
        want = numpy_voxel_mip(w)
        nxyz = [tuple(int(v) for v in p[:3]) for p in w.params.reshape(-1, 4)]
        for k, (nx, ny, nz) in enumerate(nxyz):
            Bk = mip_block()
            used = ((nx + Bk - 1) // Bk) * ((ny + Bk - 1) // Bk) * ((nz + Bk - 1) // Bk)
            assert np.array_equal(mip[k, :used].cpu().numpy().view(np.uint16), want[k, :used])
    rm = load_robot("g1_29")
    B = 48
    q = humanoid_q(rm, B, seed=91)[:, None, :]
    env = (np.arange(B) % 2).astype(np.int32)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0))
    outs = []
    for use in (True, False):
        eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV), use_voxel_mip=use)
        assert (eng._vs.mip is not None and eng._vs.mip != 0) == use
        o = eng.evaluate_action(T(q), env_query_idx=T(env))
        outs.append((o.cost.clone(), o.grad_q.clone(), o.scene_cost.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b)
    assert float(outs[0][2].sum()) > 0
    # per-op kernel with the level attached to the data holder
    sph = Kinematics(rm, DEV).compute_kinematics(T(q)).robot_spheres.detach()
    res = []
    for use in (False, True):
        vd = VoxelData.from_world(vox, DEV)
        if use:
            vd.build_mip()
        buf = CollisionBuffer.from_shape(tuple(sph.shape), DEV)
        d = SphereObstacleCollision.apply(sph, buf, SceneData(None, vd), T(np.array([5000.0], np.float32)),
                                          T(np.array([0.02], np.float32)), None, T(env), True, False)
        res.append((d.clone(), buf.gradient.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])


# ------------------------------------------------------------------------------------------------
# the shipped MPC cost (content/configs/task/mpc/lbfgs_mpc.yml): c-space target term at weight 1000 with the 0.05
# non-terminal factor, un-retimed bound weights, swept collision + speed metric, pose cost on every waypoint
# ------------------------------------------------------------------------------------------------
def _mpc_inputs(rm, B, H, seed):
    q = random_walk_q(rm, B, H, seed=seed)
    rng = np.random.default_rng(seed)
    dt = np.full(B, 0.05, np.float32)
    v = (np.gradient(q, axis=1).astype(np.float32) / 0.05) if H > 1 else np.zeros_like(q)
    a_ = rng.normal(0, 5.0, size=q.shape).astype(np.float32)
    j_ = rng.normal(0, 200.0, size=q.shape).astype(np.float32)
    target = random_q(rm, 3, seed=seed + 1, scale=0.5)
    tidx = (np.arange(B) % 3).astype(np.int32)
    dofw = np.linspace(0.5, 1.5, rm.num_dof).astype(np.float32)
    dofw[2] = 0.0                                            # a dof the target ignores
    return q, v, a_, j_, dt, target, tidx, dofw


@pytest.mark.parametrize("B,H", [(5, 30), (3, 7), (4, 1)])
@pytest.mark.parametrize("with_dofw", [True, False])
def test_mpc_config_rollout_vs_oracle(B, H, with_dofw):
    rm = load_robot("franka")
    q, v, a_, j_, dt, target, tidx, dofw = _mpc_inputs(rm, B, H, 90 + H)
    if not with_dofw:
        dofw = None
    cfg = RolloutConfig.mpc()
    assert cfg.cspace_target_weight == 1000.0 and cfg.cspace_non_terminal_weight_factor == 0.05 and not cfg.retime_weights
    cub, vox = make_benchmark_cuboid_world(), small_voxel_world()
    gp, gq = goal_from_q(rm, random_q(rm, B, seed=92))
    idx = np.arange(B, dtype=np.int32)
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    eng.update_goal(T(gp), T(gq), T(idx))
    with pytest.raises(ValueError):
        eng.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))   # target weight > 0 but no target given
    eng.update_cspace_target(T(target), T(tidx), None if dofw is None else T(dofw))
    out = eng.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt))
    torch.cuda.synchronize()
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), world_cuboid=cub, world_voxel=vox, goal_pos=gp, goal_quat=gq,
                               idxs_goal=idx, vel=v, acc=a_, jerk=j_, dt=dt, cspace_target=target, idxs_cspace_target=tidx,
                               cspace_target_dof_weight=dofw)
    cost_close(out.cspace_cost.cpu().numpy(), want["cspace_cost"], rtol=2e-4)
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    for got, wantg in zip((out.grad_vel, out.grad_acc, out.grad_jerk), want["cspace_grads"][1:4]):
        grad_close(got.cpu().numpy(), wantg)
    # the target term is live and the non-terminal factor applies: without it the c-space cost is smaller, and the terminal
    # waypoint carries 1 / 0.05 of a non-terminal waypoint's weight
    cfg0 = RolloutConfig(**{**cfg.__dict__, "cspace_target_weight": 0.0})
    eng0 = RolloutEngine(rm, cfg0, DEV, CuboidData.from_world(cub, DEV), VoxelData.from_world(vox, DEV))
    eng0.update_goal(T(gp), T(gq), T(idx))
    c0 = eng0.evaluate_action(T(q), vel=T(v), acc=T(a_), jerk=T(j_), dt=T(dt)).cspace_cost.clone()
    diff = (out.cspace_cost - c0).cpu().numpy()                         # = tw_h * dofw_d * err^2
    dw = np.ones(rm.num_dof, np.float32) if dofw is None else dofw
    err2 = (q - target[tidx][:, None, :]) ** 2
    tw = np.full(H, 1000.0 * 0.05, np.float32)
    tw[-1] = 1000.0
    np.testing.assert_allclose(diff, tw[None, :, None] * dw[None, None, :] * err2, rtol=2e-3, atol=2e-3 * np.abs(diff).max())


def test_ik_position_cspace_target_vs_oracle():
    """POSITION c-space cost with a live target term (wp_cspace_position.py target block; retract configuration of IK)."""
    rm = load_robot("franka")
    B = 40
    q = random_q(rm, B, seed=95)[:, None, :]
    cfg = RolloutConfig.ik()
    cfg.cspace_target_weight = 25.0
    target = random_q(rm, 2, seed=96, scale=0.3)
    tidx = (np.arange(B) % 2).astype(np.int32)
    dofw = np.array([1, 1, 0, 2, 1, 0.5, 1], np.float32)
    gp, gq = goal_from_q(rm, random_q(rm, 4, seed=97))
    idx = (np.arange(B) % 4).astype(np.int32)
    cub = make_benchmark_cuboid_world()
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV))
    eng.update_goal(T(gp), T(gq), T(idx))
    eng.update_cspace_target(T(target), T(tidx), T(dofw))
    out = eng.evaluate_action(T(q))
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), world_cuboid=cub, goal_pos=gp, goal_quat=gq, idxs_goal=idx,
                               cspace_target=target, idxs_cspace_target=tidx, cspace_target_dof_weight=dofw)
    cost_close(out.cspace_cost.cpu().numpy(), want["cspace_cost"], rtol=2e-4)
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    assert float(out.cspace_cost[:, :, 2].abs().max()) == float(T(want["cspace_cost"])[:, :, 2].abs().max())


def test_mpc_full_size_sampled_seeds_vs_oracle():
    """BASELINE config 4 as the reference defines it: 1024 particles x 30 waypoints, shipped MPC weights, 256^3 ESDF;
    whole trajectories of sampled seeds against the oracle, plus determinism and a CUDA-graph replay."""
    rm = load_robot("franka")
    B, H = 1024, 30
    sdf = make_box_esdf(n=256, voxel_size=0.01, num_boxes=12, seed=0, xp=torch)
    vox = VoxelData(T(np.array([[[256, 256, 256, 0.01]]], np.float32)), T(np.array([[[0, 0, 0, 1, 0, 0, 0, 0]]], np.float32)),
                    torch.ones((1, 1), dtype=torch.uint8, device=DEV), torch.ones(1, dtype=torch.int32, device=DEV),
                    sdf.reshape(1, 1, -1).contiguous(), 1, 1, 100.0)
    q, v, a_, j_, dt, target, tidx, dofw = _mpc_inputs(rm, B, H, 98)
    cfg = RolloutConfig.mpc()
    gp, gq = goal_from_q(rm, random_q(rm, 8, seed=99))
    idx = (np.arange(B) % 8).astype(np.int32)
    eng = RolloutEngine(rm, cfg, DEV, voxel=vox)
    eng.update_goal(T(gp), T(gq), T(idx))
    eng.update_cspace_target(T(target), T(tidx), T(dofw))
    tq, tv, ta, tj, tdt = T(q), T(v), T(a_), T(j_), T(dt)
    o = eng.evaluate_action(tq, vel=tv, acc=ta, jerk=tj, dt=tdt)
    c1, g1 = o.cost.clone(), o.grad_q.clone()
    o = eng.evaluate_action(tq, vel=tv, acc=ta, jerk=tj, dt=tdt)
    assert torch.equal(c1, o.cost) and torch.equal(g1, o.grad_q) and torch.isfinite(g1).all()
    sel = np.arange(0, B, 128)
    vw = VoxelWorld(vox.params.cpu().numpy(), vox.inv_pose.cpu().numpy(), np.ones((1, 1), np.uint8), np.ones(1, np.int32),
                    sdf.reshape(1, 1, -1).cpu().numpy(), 100.0)
    want = O.rollout_cost_grad(rm, q[sel], cfg.to_oracle_cfg(1), world_voxel=vw, goal_pos=gp, goal_quat=gq, idxs_goal=idx[sel],
                               vel=v[sel], acc=a_[sel], jerk=j_[sel], dt=dt[sel], cspace_target=target,
                               idxs_cspace_target=tidx[sel], cspace_target_dof_weight=dofw)
    np.testing.assert_allclose(c1.cpu().numpy()[sel], want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(g1.cpu().numpy()[sel], want["grad_q"], rtol=2e-3, scale=2e-5)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        eng.evaluate_action(tq, vel=tv, acc=ta, jerk=tj, dt=tdt)
    eng.out.cost.zero_()
    eng.out.grad_q.zero_()
    gr.replay()
    torch.cuda.synchronize()
    assert torch.equal(eng.out.cost, c1) and torch.equal(eng.out.grad_q, g1)


@pytest.mark.parametrize("mode", ["discrete", "swept"])
def test_fused_rollout_sphere_configs(mode):
    """Several link-sphere configurations (attached objects per environment, kinematics_forward_helper.cuh:232-233): a row
    uses robot_spheres[env_query_idx[b]]; config 1 grows / moves some spheres and disables others."""
    import copy
    rm = copy.copy(load_robot("franka"))
    ls0 = np.asarray(rm.link_spheres, np.float32).reshape(-1, 4)
    ls1 = ls0.copy()
    ls1[-4:, 3] = 0.08                                   # an attached object: big spheres on the last link
    ls1[-4:, :3] += np.array([0.0, 0.0, 0.12], np.float32)
    ls1[5:8, 3] = -1.0                                   # disabled spheres
    ls2 = ls0.copy()
    ls2[:, 3] = np.where(ls0[:, 3] >= 0, ls0[:, 3] * 1.3, ls0[:, 3])
    rm.link_spheres = np.stack([ls0, ls1, ls2])
    B, H = 12, (1 if mode == "discrete" else 5)
    q = random_walk_q(rm, B, H, seed=101) if H > 1 else random_q(rm, B, seed=101)[:, None, :]
    cub, vox = _two_env_worlds()
    cub3 = type(cub)(np.concatenate([cub.dims, cub.dims[:1]]), np.concatenate([cub.inv_pose, cub.inv_pose[:1]]),
                     np.concatenate([cub.enable, cub.enable[:1]]), np.concatenate([cub.count, cub.count[:1]]))
    env = (np.arange(B) % 3).astype(np.int32)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0),
                        use_sweep=(mode == "swept"), use_speed_metric=False)
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub3, DEV), store_fk_outputs=True)
    out = eng.evaluate_action(T(q), env_query_idx=T(env))
    torch.cuda.synchronize()
    want = O.rollout_cost_grad(rm, q, cfg.to_oracle_cfg(1), world_cuboid=cub3, env_query_idx=env)
    np.testing.assert_allclose(out.robot_spheres.cpu().numpy(), want["spheres"], atol=1e-5)
    cost_close(out.self_cost.cpu().numpy(), want["self_cost"])
    np.testing.assert_allclose(out.scene_cost.cpu().numpy(), want["scene_cost"], rtol=2e-4,
                               atol=1e-5 * max(want["scene_cost"].max(), 1e-6))
    np.testing.assert_allclose(out.cost.cpu().numpy(), want["cost_bh"], rtol=2e-4, atol=1e-5 * want["cost_bh"].max())
    grad_close(out.grad_q.cpu().numpy(), want["grad_q"], rtol=2e-3, scale=2e-5)
    assert np.abs(want["spheres"][1, :, -4:, 3] - 0.08).max() < 1e-6 and (want["spheres"][1, :, 5:8, 3] < 0).all()
    assert want["scene_cost"].sum() + want["self_cost"].sum() > 0


@pytest.mark.parametrize("robot,n,buried", [("g1_29", 24, False), ("g1_29", 10, True), ("g1_43", 8, True), ("franka", 32, True)])
def test_big_robot_kernel_matches_standard_kernel_and_oracle(monkeypatch, robot, n, buried):
    """rollout_fused_big_kernel (gradient list instead of the dense sphere-gradient array, padded radii rebuilt on the fly,
    ticket-counter row queue) against rollout_fused_kernel on the same rows and against the oracle.  `buried`: the world is solid
    around the robot, every sphere collides, the list overflows and the row finishes through the dense force / torque
    accumulators -- both J^T forms of the new kernel are covered.  CB200_BIG forces the variant (default: big for humanoids)."""
    from curobo_b200.world import VoxelWorld
    rm = load_robot(robot)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0), pose_weight=(2000.0, 100.0))
    q = (humanoid_q(rm, n, seed=61) if robot != "franka" else random_q(rm, n, seed=61))[:, None, :]
    gp, gq = goal_from_q(rm, (humanoid_q(rm, 2, seed=62, scale=0.5) if robot != "franka" else random_q(rm, 2, seed=62)))
    idx = (np.arange(n) % 2).astype(np.int32)
    if buried:
        g = np.stack(np.meshgrid(*[np.arange(48)] * 3, indexing="ij"), -1).astype(np.float32)
        sdf = (np.linalg.norm(g - 23.5, axis=-1) * 0.06 - 1.2).astype(np.float32)      # a solid ball of radius 1.2 m around the base
        vox = VoxelWorld.from_grid(sdf, 0.06)
    else:
        vox = small_voxel_world()
    outs = {}
    monkeypatch.setenv("CB200_TEAM", "0")          # (few rows would select the team variant, which has its own test below)
    for flag in ("0", "1"):
        monkeypatch.setenv("CB200_BIG", flag)
        out, want = check_against_oracle(rm, cfg, q, vox=vox, goal=(gp, gq), idx=idx)
        outs[flag] = (out.cost.clone(), out.grad_q.clone(), out.scene_cost.clone(), out.self_cost.clone())
    if buried:
        assert int((want["scene_cost"] > 0).sum(-1).min()) > (96 if robot != "franka" else 40), "every row must overflow the list"
    a, b = outs["0"], outs["1"]
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])          # per-term costs: same arithmetic, bit for bit
    assert torch.allclose(a[0], b[0], rtol=1e-6, atol=0.0)
    assert torch.allclose(a[1], b[1], rtol=2e-4, atol=2e-6 * float(a[1].abs().max()))     # different summation order in J^T
    # the ticket counter is re-armed by every launch
    eng = RolloutEngine(rm, cfg, DEV, None, VoxelData.from_world(vox, DEV))
    eng.update_goal(T(gp), T(gq), T(idx))
    first = eng.evaluate_action(T(q)).grad_q.clone()
    for _ in range(3):
        again = eng.evaluate_action(T(q)).grad_q
    torch.cuda.synchronize()
    assert torch.equal(first, again) and int(eng._work_counter.abs().sum()) == 0


@pytest.mark.parametrize("robot,n,team,scene", [("g1_29", 24, 2, "esdf"), ("g1_29", 9, 4, "esdf"), ("g1_43", 8, 2, "esdf"),
                                                 ("g1_29", 10, 2, "buried"), ("franka", 32, 4, "cuboid"),
                                                 ("g1_29", 12, 4, "both")])
def test_team_kernel_matches_big_kernel_and_oracle(monkeypatch, robot, n, team, scene):
    """rollout_fused_team_kernel (TEAM warps share one row: strided spheres / link pairs / list segments, partial J^T sums,
    named barriers) against rollout_fused_big_kernel on the same rows and against the oracle.  "buried": every sphere collides, a
    list segment overflows and the team's first warp redoes the row through the single-warp code.  CB200_TEAM forces the team
    size (default: by rows vs resident warps)."""
    from curobo_b200.world import VoxelWorld
    rm = load_robot(robot)
    cfg = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                        cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0), pose_weight=(2000.0, 100.0))
    q = (humanoid_q(rm, n, seed=71) if robot != "franka" else random_q(rm, n, seed=71))[:, None, :]
    gp, gq = goal_from_q(rm, (humanoid_q(rm, 2, seed=72, scale=0.5) if robot != "franka" else random_q(rm, 2, seed=72)))
    idx = (np.arange(n) % 2).astype(np.int32)
    vox = cub = None
    if scene == "buried":
        g = np.stack(np.meshgrid(*[np.arange(48)] * 3, indexing="ij"), -1).astype(np.float32)
        vox = VoxelWorld.from_grid((np.linalg.norm(g - 23.5, axis=-1) * 0.06 - 1.2).astype(np.float32), 0.06)
    if scene in ("esdf", "both"):
        vox = small_voxel_world()
    if scene in ("cuboid", "both"):
        cub = make_benchmark_cuboid_world()
    monkeypatch.setenv("CB200_BIG", "1")
    outs = {}
    for flag in ("0", str(team)):
        monkeypatch.setenv("CB200_TEAM", flag)
        out, want = check_against_oracle(rm, cfg, q, vox=vox, cub=cub, goal=(gp, gq), idx=idx)
        assert last_variant() == {"0": 4, "2": 5, "4": 6}[flag]      # CB200_VARIANT_BIG / TEAM2 / TEAM4
        outs[flag] = (out.cost.clone(), out.grad_q.clone(), out.scene_cost.clone(), out.self_cost.clone())
    if scene == "buried":
        assert int((want["scene_cost"] > 0).sum(-1).min()) > 96, "every row must overflow its list segments"
    else:
        assert float(want["scene_cost"].sum()) > 0 and (robot == "franka" or float(want["self_cost"].sum()) > 0)
    a, b = outs["0"], outs[str(team)]
    assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3])          # per-term costs: same arithmetic, bit for bit
    assert torch.allclose(a[0], b[0], rtol=2e-6, atol=0.0)
    assert torch.allclose(a[1], b[1], rtol=2e-4, atol=2e-6 * float(a[1].abs().max()))     # partial sums per warp of the team
    monkeypatch.delenv("CB200_BIG")
    monkeypatch.delenv("CB200_TEAM")                                    # default selection: few rows -> a team kernel
    eng = RolloutEngine(rm, cfg, DEV, CuboidData.from_world(cub, DEV) if cub is not None else None,
                        VoxelData.from_world(vox, DEV) if vox is not None else None)
    eng.update_goal(T(gp), T(gq), T(idx))
    first = eng.evaluate_action(T(q)).grad_q.clone()
    for _ in range(3):
        again = eng.evaluate_action(T(q)).grad_q
    torch.cuda.synchronize()
    assert torch.equal(first, again) and int(eng._work_counter.abs().sum()) == 0
