"""Pin the oracle: the reference's own golden vectors / known-answer tests for this path."""
import os

import numpy as np
import pytest

from helpers import random_q
from curobo_b200.robot_model import load_robot
from oracle import rollout_oracle as O
from voxel_cases import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_fk_golden_vector_franka():
    """curobo/tests/_src/robot/kinematics/test_kinematics.py:57-82"""
    rm = load_robot("franka")
    q = np.array([[0.0, -1.2, 0.0, -2.0, 0.0, 1.0, 0.0]], np.float32)
    _, sph, pos, quat = O.fk_forward(rm, q)
    assert rm.tool_frames == ["panda_hand"]
    np.testing.assert_allclose(pos[0, 0], [6.0860e-02, -4.7547e-12, 7.6373e-01], atol=1e-5)
    np.testing.assert_allclose(quat[0, 0], [0.0382, 0.9193, 0.3808, 0.0922], atol=1e-4)
    # attached-object spheres sit on the hand frame: x ~ 0.061, y ~ 0 (:98-99)
    assert abs(sph[0, -1, 0] - 0.061) < 1e-3 and abs(sph[0, -1, 1]) < 1e-5
    assert sph[0, -1, 3] == -100.0


def test_robot_fixture_sizes():
    """Pair counts / sizes quoted in SURVEY.md section 8 (exact from collision_spheres x ignore lists)."""
    exp = {"franka": (7, 65, 1, 818, 1), "g1_29": (35, 400, 14, 55414, 1), "g1_43": (49, 674, 4, 162111, 2)}
    for name, (D, S, L, P, nblk) in exp.items():
        rm = load_robot(name)
        assert (rm.num_dof, rm.num_spheres, rm.num_tool_frames, rm.collision_pairs.shape[0], rm.num_blocks_per_batch) == (D, S, L, P, nblk)
        assert (rm.collision_pairs[:, 0] < rm.collision_pairs[:, 1]).all()
        assert (rm.link_map[1:] < np.arange(1, rm.num_links)).all()


def test_self_collision_default_pose_small():
    """tests/_src/cost/test_cost_self_collision.py:171-199: default Franka pose cost <= 0.1 (w = 1)."""
    rm = load_robot("franka")
    _, sph, _, _ = O.fk_forward(rm, rm.default_joint_position[None])
    c, g, k = O.self_collision(sph, rm.sphere_padding, rm.collision_pairs, 1.0)
    assert c[0] <= 0.1
    # lock-joint invariance analogue (:182-193): locked finger joints are baked into fixed transforms
    assert rm.num_dof == 7


def test_self_collision_folded_arm_collides():
    rm = load_robot("franka")
    q = np.array([[0.0, 1.7, 0.0, -3.0, 0.0, 3.7, 0.0]], np.float32)
    _, sph, _, _ = O.fk_forward(rm, q)
    c, g, k = O.self_collision(sph, rm.sphere_padding, rm.collision_pairs, 2.0)
    assert c[0] > 0 and k[0] >= 0
    i, j = rm.collision_pairs[k[0]]
    np.testing.assert_allclose(g[0, i, :3], 2.0 * (sph[0, j, :3] - sph[0, i, :3]), rtol=1e-6)
    np.testing.assert_allclose(g[0, j, :3], -g[0, i, :3])
    assert g[0, i, 3] == -2.0 and g[0, j, 3] == -2.0
    assert np.count_nonzero(np.abs(g[0]).sum(-1)) == 2


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_voxel_property_cases_oracle(case):
    name, world, spheres, eta, sweep, check = case
    if sweep == "both":
        c0, g0 = O.scene_collision(spheres[:, :1], 1.0, eta, None, world)
        c1, g1 = O.scene_collision(spheres, 1.0, eta, None, world, sweep=True)
        np.testing.assert_allclose(c1, np.broadcast_to(c0, c1.shape), rtol=1e-6)
        np.testing.assert_allclose(g1, np.broadcast_to(g0, g1.shape), rtol=1e-6, atol=1e-7)
        assert c0.max() > 0
        return
    c, g = O.scene_collision(spheres, 1.0, eta, None, world, sweep=bool(sweep))
    assert check(c, g), (name, c, g)


@pytest.mark.parametrize("robot", ["franka", "g1_29"])
def test_fk_backward_matches_finite_differences(robot):
    """J^T backward vs central differences of a random linear functional of spheres + tool positions
    (tolerances as the reference's FD grad-checks, tests/_src/robot/kinematics/test_jacobian_gradcheck.py:144-220)."""
    rm = load_robot(robot)
    rng = np.random.default_rng(0)
    q = random_q(rm, 3, seed=5, scale=0.5).astype(np.float64)
    ws = rng.normal(size=(rm.num_spheres, 3)).astype(np.float32)
    wp = rng.normal(size=(rm.num_tool_frames, 3)).astype(np.float32)

    def f(qq):
        _, sph, pos, _ = O.fk_forward(rm, qq.astype(np.float32))
        return (sph[..., :3].astype(np.float64) * ws).sum((1, 2)) + (pos.astype(np.float64) * wp).sum((1, 2))

    cum, sph, pos, quat = O.fk_forward(rm, q.astype(np.float32))
    gs = np.zeros_like(sph)
    gs[..., :3] = ws
    gq = O.fk_backward(rm, cum, gs, np.broadcast_to(wp, pos.shape), np.zeros_like(quat))
    eps = 1e-3
    fd = np.zeros_like(q)
    for d in range(rm.num_dof):
        dq = np.zeros_like(q)
        dq[:, d] = eps
        fd[:, d] = (f(q + dq) - f(q - dq)) / (2 * eps)
    np.testing.assert_allclose(gq, fd, rtol=5e-3, atol=5e-3 * np.abs(fd).max())


def test_fk_backward_orientation_convention():
    """The reference's quaternion 'gradient' is a convention, not d/dquat: the tool-pose cost emits
    g_quat = q (x) (omega, 0) (cost/wp_tool_pose.py:113-126) and FK backward maps it back with
    omega' = 1/2 E(q)^T g_quat (quaternion_util.cuh:86-103) = omega/2 for unit q, which is then dotted
    with the WORLD-frame joint axes (kinematics_joint_util.cuh:49-62).  Check that composition."""
    rm = load_robot("franka")
    q = random_q(rm, 5, seed=2, scale=0.5)
    cum, sph, pos, quat = O.fk_forward(rm, q)
    om = np.random.default_rng(1).normal(size=(5, 1, 3)).astype(np.float32)
    cq = np.concatenate([quat[..., 1:], quat[..., :1]], -1)
    rate = O._quat_mul_xyzw(cq, np.concatenate([om, np.zeros_like(om[..., :1])], -1))
    g_quat = np.concatenate([rate[..., 3:], rate[..., :3]], -1)
    np.testing.assert_allclose(O.quat_grad_to_omega(quat, g_quat), 0.5 * om, rtol=1e-5, atol=1e-6)
    gq = O.fk_backward(rm, cum, None, np.zeros_like(pos), g_quat)
    exp = np.zeros_like(gq)
    for j in range(rm.num_links):
        if rm.joint_map_type[j] >= 3 and j in rm.link_chain_data[rm.link_chain_offsets[rm.tool_frame_map[0]]:rm.link_chain_offsets[rm.tool_frame_map[0] + 1]]:
            axis = cum[:, j, :, rm.joint_map_type[j] - 3]
            exp[:, rm.joint_map[j]] += rm.joint_offset_map[j, 0] * np.sum(axis * 0.5 * om[:, 0], -1)
    np.testing.assert_allclose(gq, exp, rtol=1e-4, atol=1e-6)


def test_golden_fixtures_match_oracle():
    """The committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py) are what the
    GPU parity tests compare against; guard against the oracle drifting away from them."""
    from golden.make_golden import build_cases
    for name, arrays in build_cases().items():
        ref = np.load(os.path.join(GOLD, name + ".npz"))
        for k, v in arrays.items():
            np.testing.assert_allclose(ref[k], v, rtol=1e-5, atol=1e-6, err_msg=f"{name}:{k}")
