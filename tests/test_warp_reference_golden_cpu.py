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


@pytest.mark.parametrize("name", ["collision_discrete", "collision_multi_env", "collision_swept", "collision_swept_speed",
                                  "collision_edge_discrete", "collision_edge_swept", "collision_edge_swept_speed"])
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


@pytest.mark.parametrize("name", ["collision_discrete", "collision_swept", "collision_swept_speed", "collision_edge_discrete",
                                  "collision_edge_swept", "collision_edge_swept_speed"])
def test_product_scene_math_compiled_for_the_host_matches_the_reference_source(name):
    """No oracle in between: the product's __host__ __device__ scene-collision arithmetic (curobo_b200/csrc/cb200_math.cuh,
    compiled for the host by tests/hostmath) against the reference-source fixture, including the exact ESDF cull level."""
    from helpers import hm_scene, hostmath, numpy_voxel_mip
    c = case(name)
    cub, vox = worlds(c)
    speed = "speed_dt" in c
    for mip in (None, numpy_voxel_mip(vox)):
        cost, grad = hm_scene(hostmath(), c["spheres"], float(c["weight"]), float(c["eta"]), "swept" in name, speed,
                              float(c["speed_dt"]) if speed else 0.0, cub=cub, vox=vox, mip=mip)
        close(cost, c["cost"], 1e-4, "cost")
        close(grad, c["grad"], 5e-4, "gradient")
        assert np.array_equal(cost > 0, c["cost"] > 0)


def test_warp_stand_in_semantics():
    """The stand-in's own rules: C integer division, float32 arithmetic, xyzw quaternions, transform algebra."""
    import sys
    shim = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "warp_shim")
    sys.path.insert(0, shim)
    try:
        import warp as wp
        assert wp.int32(7) / wp.int32(2) == 3 and isinstance(wp.int32(7) / 2, wp.int32) and (-wp.int32(7)) / 2 == -3
        assert 9 / wp.int32(2) == 4 and wp.int32(7) % 4 == 3 and wp.int32(2.9) == 2
        x = wp.float32(0.1) * 3.0
        assert isinstance(x, np.float32)
        assert wp.sign(wp.float32(0.0)) == 1 and wp.sign(wp.float32(-2.0)) == -1
        q = wp.quat(0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4))           # 90 degrees about z
        v = wp.quat_rotate(q, wp.vec3(1.0, 0.0, 0.0))
        assert np.allclose(v.v, [0, 1, 0], atol=1e-6)
        t = wp.transform(wp.vec3(1.0, 2.0, 3.0), q)
        p = wp.transform_point(t, wp.vec3(1.0, 0.0, 0.0))
        assert np.allclose(p.v, [1, 3, 3], atol=1e-6)
        back = wp.transform_point(wp.transform_inverse(t), p)
        assert np.allclose(back.v, [1, 0, 0], atol=1e-6)
        assert np.allclose((q * wp.quat_inverse(q)).v, [0, 0, 0, 1], atol=1e-6)
        s = wp.float32(2.0) * wp.vec3(1.0, 2.0, 3.0)                          # numpy scalar on the left defers to the vector
        assert isinstance(s, wp.vec3) and np.allclose(s.v, [2, 4, 6])
        a = wp.from_numpy(np.zeros(4, np.float32))
        wp.atomic_add(a, 2, wp.float32(1.5))
        wp.atomic_add(a, 2, wp.float32(1.0))
        assert a.data[2] == 2.5
        seen = []

        @wp.kernel
        def k(out: wp.array(dtype=wp.int32), n: wp.int32):
            t_ = wp.tid()
            out[t_] = t_ / n
            seen.append(type(t_))

        o = wp.from_numpy(np.zeros(6, np.int32), dtype=wp.int32)
        wp.launch(k, dim=6, inputs=[o, 4])
        assert list(o.data) == [0, 0, 0, 0, 1, 1] and seen[0] is wp.int32
    finally:
        sys.path.remove(shim)
        sys.modules.pop("warp", None)
        for n in [n for n in sys.modules if n.startswith("warp.")]:
            sys.modules.pop(n)


@pytest.mark.parametrize("method", [0, 1])
def test_product_tool_pose_math_compiled_for_the_host_matches_the_reference_source(method):
    """The product's tool_pose_cost (cb200_math.cuh, host build) per (batch, horizon, link) against the reference-source fixture."""
    import ctypes as C
    from helpers import hostmath, ptr
    c = case(f"tool_pose_method{method}")
    lib = hostmath()
    B, H, L, _ = c["pos"].shape
    NG = c["goal_pos"].shape[2]
    for b in range(B):
        for h in range(H):
            term = (h == H - 1) or H == 1
            for l in range(L):
                axes = np.ascontiguousarray((c["axes_t"] if term else c["axes_nt"])[l], np.float32)
                tol = (c["tol_t"] if term else c["tol_nt"])[l]
                g = int(c["idxs_goal"][b])
                out = np.zeros(12, np.float32)
                idx = C.c_int(-1)
                lib.hm_tool_pose(ptr(np.ascontiguousarray(c["pos"][b, h, l])), ptr(np.ascontiguousarray(c["quat"][b, h, l])),
                                 ptr(np.ascontiguousarray(c["goal_pos"][g, l])), ptr(np.ascontiguousarray(c["goal_quat"][g, l])),
                                 C.c_int(NG), C.c_float(c["weight"][0]), C.c_float(c["weight"][1]), ptr(axes), C.c_float(tol[0]),
                                 C.c_float(tol[1]), C.c_int(method), ptr(out), C.byref(idx))
                assert idx.value == c["goalset_idx"][b, h, l]
                sc = max(float(np.abs(c["distance"]).max()), 1.0)
                assert np.allclose(out[0:2], c["distance"][b, h, 2 * l:2 * l + 2], rtol=2e-4, atol=1e-5 * sc)
                assert np.allclose(out[2], c["pos_dist"][b, h, l], rtol=2e-4, atol=1e-5)
                assert np.allclose(out[3], c["rot_dist"][b, h, l], rtol=2e-4, atol=1e-5)
                assert np.allclose(out[4:7], c["grad_pos"][b, h, l], rtol=5e-4, atol=1e-4 * float(np.abs(c["grad_pos"]).max()))
                assert np.allclose(out[7:11], c["grad_quat"][b, h, l], rtol=2e-3, atol=2e-4 * float(np.abs(c["grad_quat"]).max()))
