"""The oracle (oracle/rollout_oracle.py) against outputs of the REFERENCE's own Warp kernel sources.

tests/golden/warp_reference_golden.npz was produced by tests/golden/make_warp_golden.py: the reference's `@wp.kernel` /
`@wp.func` Python bodies (geom/collision/wp_collision_kernel.py, wp_sweep_collision_kernel.py, wp_speed_metric.py,
geom/data/data_cuboid.py, data_voxel.py, cost/wp_tool_pose.py, cost/wp_cspace_state.py, cost/wp_cspace_position.py), imported
from the reference tree and executed on the CPU one thread at a time under a pure-Python stand-in for Warp's builtins
(oracle/warp_shim).  This pins the parts of the hot path the reference implements in Warp -- which cannot be compiled or run
here -- on the reference's own source rather than on a restatement; what remains restated is the handful of Warp builtins
(quaternion / transform algebra, integer division), documented in the stand-in.  The GPU tests then hold the kernels to the
oracle on the same kinds of inputs."""
import os

import numpy as np
import pytest

from curobo_b200.world import CuboidWorld, VoxelWorld
from oracle import rollout_oracle as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "warp_reference_golden.npz"))


def case(name):
    pre = name + "/"
    return {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}


def close(got, want, rtol=2e-5, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert np.allclose(got, want, rtol=rtol, atol=rtol * max(float(np.abs(want).max()), 1e-30)), \
        (what, float(np.abs(got - want).max()), float(np.abs(want).max()))


def worlds(c):
    cub = CuboidWorld(c["cub_dims"], c["cub_inv_pose"], c["cub_enable"], c["cub_count"]) if "cub_dims" in c else None
    vox = VoxelWorld(c["vox_params"], c["vox_inv_pose"], c["vox_enable"], c["vox_count"], c["vox_features"],
                     float(c["vox_max_dist"])) if "vox_params" in c else None
    return cub, vox


@pytest.mark.parametrize("name", ["collision_discrete", "collision_multi_env", "collision_swept", "collision_swept_speed"])
def test_scene_collision_matches_the_reference_source(name):
    c = case(name)
    cub, vox = worlds(c)
    cost, grad = O.scene_collision(c["spheres"], float(c["weight"]), float(c["eta"]), world_cuboid=cub, world_voxel=vox,
                                   env_query_idx=c.get("env_query_idx"), sweep="swept" in name,
                                   speed_dt=float(c["speed_dt"]) if "speed_dt" in c else None)
    assert (c["cost"] > 0).sum() >= 5, "the case must contain collisions"
    close(cost, c["cost"], 5e-5, "cost")
    close(grad, c["grad"], 2e-4, "gradient")
    assert np.array_equal(cost > 0, c["cost"] > 0), "a different set of spheres is in collision"


@pytest.mark.parametrize("method", [0, 1])
def test_tool_pose_matches_the_reference_source(method):
    c = case(f"tool_pose_method{method}")
    cost, gp, gq, gi, pe, re = O.tool_pose_cost(c["pos"], c["quat"], c["goal_pos"], c["goal_quat"], c["idxs_goal"], c["weight"],
                                                c["axes_t"], c["axes_nt"], c["tol_t"], c["tol_nt"], use_lie_group=bool(method))
    assert np.array_equal(gi, c["goalset_idx"])
    close(cost, c["distance"], 5e-5, "distance")
    close(pe, c["pos_dist"], 5e-5, "position distance")
    close(re, c["rot_dist"], 5e-5, "rotation distance")
    close(gp, c["grad_pos"], 1e-4, "position gradient")
    close(gq, c["grad_quat"], 5e-4, "quaternion-rate gradient")


@pytest.mark.parametrize("retime", [0, 1])
def test_cspace_state_cost_matches_the_reference_source(retime):
    c = case(f"cspace_state_retime{retime}")
    lim = dict(p=c["lim_p"], v=c["lim_v"], a=c["lim_a"], j=c["lim_j"], tau=c["lim_tau"])
    cost, g = O.cspace_state_cost(c["q"], c["v"], c["a"], c["j"], c["dt"], lim, c["weight"], c["act"], c["reg"], bool(retime),
                                  bool(retime), effort=c["tau"], target=c["target"], idxs_target=c["idxs_target"],
                                  target_weight=float(c["target_weight"][0]), non_terminal_factor=float(c["ntf"][0]),
                                  target_dof_weight=c["dof_weight"])
    close(cost, c["cost"], 1e-6, "cost")
    for got, n in zip(g, "pvajt"):
        close(got, c[f"grad_{n}"], 1e-6, f"grad_{n}")
    assert np.abs(c["grad_t"]).max() > 0, "the effort channel must be live"


def test_cspace_position_cost_matches_the_reference_source():
    c = case("cspace_position")
    cost, g = O.cspace_position_cost(c["q"], c["lim_p"], c["weight"], c["act"], target=c["target"], idxs_target=c["idxs_target"],
                                     target_weight=float(c["target_weight"][0]), target_dof_weight=c["dof_weight"])
    close(cost, c["cost"], 1e-6, "cost")
    close(g, c["grad_p"], 1e-6, "grad_p")
