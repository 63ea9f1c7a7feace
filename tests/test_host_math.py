"""CPU unit tests of the kernels' scalar building blocks (curobo_b200/csrc/cb200_math.cuh compiled for the
host into a TEST-ONLY library) against the oracle.  Two independent statements of the reference's
arithmetic (C++ and numpy) must agree before any GPU time is spent."""
import ctypes as C

import numpy as np
import pytest

from helpers import hm_scene, hostmath, ptr, small_voxel_world
from curobo_b200 import build
from curobo_b200.robot_model import load_robot
from curobo_b200.world import CuboidWorld, VoxelWorld, make_box_esdf
from oracle import rollout_oracle as O


@pytest.fixture(scope="module")
def lib():
    build.build_hostmath()
    return hostmath()


def _random_sphere_trajs(B=8, H=10, S=40, seed=0):
    rng = np.random.default_rng(seed)
    pos = rng.uniform(-1.5, 1.5, size=(B, 1, S, 3)) + np.cumsum(rng.normal(0, 0.05, size=(B, H, S, 3)), axis=1)
    rad = rng.uniform(0.01, 0.12, size=(B, 1, S, 1))
    rad[:, :, ::7] = -1.0                                     # disabled spheres are skipped
    return np.concatenate([pos, np.broadcast_to(rad, (B, H, S, 1))], -1).astype(np.float32)


def _worlds():
    cub = CuboidWorld.create([
        {"dims": [2.2, 2.2, 0.2], "pose": [0, 0, -0.1, 1, 0, 0, 0]},
        {"dims": [0.3, 0.4, 1.5], "pose": [0.45, 0.1, 0.3, 0.9238795, 0, 0.3826834, 0]},
        {"dims": [0.5, 0.5, 0.5], "pose": [-0.6, -0.5, 0.6, 0.7071068, 0.7071068, 0, 0]}], max_n=10)
    vox = small_voxel_world(n=64, voxel=0.04)
    sdf = make_box_esdf(n=48, voxel_size=0.05, num_boxes=8, seed=5)
    vox_rot = VoxelWorld.from_grid(sdf.reshape(48, 48, 48), 0.05, pose=[0.1, -0.2, 0.3, 0.9659258, 0, 0, 0.2588190])
    return {"cuboid": (cub, None), "voxel": (None, vox), "voxel_rotated": (None, vox_rot), "both": (cub, vox_rot)}


@pytest.mark.parametrize("world", ["cuboid", "voxel", "voxel_rotated", "both"])
@pytest.mark.parametrize("mode", [(0, 0), (1, 0), (1, 1)], ids=["discrete", "swept", "swept_speed"])
@pytest.mark.parametrize("eta", [0.0, 0.02])
def test_scene_collision_host_vs_oracle(lib, world, mode, eta):
    cub, vox = _worlds()[world]
    sph = _random_sphere_trajs()
    sweep, speed = mode
    c1, g1 = hm_scene(lib, sph, 5000.0, eta, sweep, speed, 0.05, cub, vox)
    c2, g2 = O.scene_collision(sph, 5000.0, eta, cub, vox, sweep=bool(sweep), speed_dt=0.05 if speed else None)
    assert (c2 > 0).sum() > 50                                 # the case actually exercises collisions
    np.testing.assert_allclose(c1, c2, rtol=1e-5, atol=1e-5 * c2.max())
    np.testing.assert_allclose(g1, g2, rtol=1e-4, atol=1e-5 * np.abs(g2).max())


def test_local_transform_and_quaternion(lib):
    rm = load_robot("g1_29")
    rng = np.random.default_rng(3)
    q = rng.uniform(-2, 2, size=(1, rm.num_dof)).astype(np.float32)
    loc = O.local_link_transforms(rm, q)[0]
    cum = O.compose_chain(rm, loc[None])[0]
    quat_ref = O.quat_from_rotation(cum[:, :, :3])
    for l in range(rm.num_links):
        jt = int(rm.joint_map_type[l])
        th = 0.0 if jt < 0 else float(np.float32(rm.joint_offset_map[l, 0]) * q[0, rm.joint_map[l]] + np.float32(rm.joint_offset_map[l, 1]))
        out = np.zeros(12, np.float32)
        f = np.ascontiguousarray(rm.fixed_transforms[l].reshape(-1))
        lib.hm_local_transform(ptr(f), C.c_int(jt), C.c_float(th), ptr(out))
        np.testing.assert_allclose(out.reshape(3, 4), loc[l], atol=2e-6)
        qt = np.zeros(4, np.float32)
        t = np.ascontiguousarray(cum[l].reshape(-1))
        lib.hm_quat_from_transform(ptr(t), ptr(qt))
        np.testing.assert_allclose(qt, quat_ref[l], atol=2e-6)
        assert qt[0] >= 0


@pytest.mark.parametrize("method", [0, 1], ids=["axis_angle", "lie"])
def test_tool_pose_host_vs_oracle(lib, method):
    rng = np.random.default_rng(7)
    n, G = 40, 3
    pos = rng.normal(size=(n, 1, 1, 3)).astype(np.float32)
    quat = rng.normal(size=(n, 1, 1, 4)).astype(np.float32)
    quat /= np.linalg.norm(quat, axis=-1, keepdims=True)
    gpos = (pos[:, 0][:, :, None, :] + rng.normal(0, 0.2, size=(n, 1, G, 3))).astype(np.float32)
    gquat = (quat[:, 0][:, :, None, :] + rng.normal(0, 0.4, size=(n, 1, G, 4))).astype(np.float32)
    gquat /= np.linalg.norm(gquat, axis=-1, keepdims=True)
    gquat[0, 0, 0] = quat[0, 0, 0]                               # exact-match goal: zero rotation error
    axes = np.array([[1.0, 0.5, 2.0, 1.0, 0.3, 1.5]], np.float32)
    tol = np.array([[1e-3, 1e-2]], np.float32)
    w = (1000.0, 30.0)
    c, gp, gq, gi, pe, re = O.tool_pose_cost(pos, quat, gpos, gquat, np.arange(n), w, axes, axes, tol, tol,
                                             use_lie_group=bool(method))
    for i in range(n):
        out = np.zeros(12, np.float32)
        idx = C.c_int(-1)
        lib.hm_tool_pose(ptr(np.ascontiguousarray(pos[i, 0, 0])), ptr(np.ascontiguousarray(quat[i, 0, 0])),
                         ptr(np.ascontiguousarray(gpos[i, 0])), ptr(np.ascontiguousarray(gquat[i, 0])), C.c_int(G),
                         C.c_float(w[0]), C.c_float(w[1]), ptr(axes[0]), C.c_float(tol[0, 0]), C.c_float(tol[0, 1]),
                         C.c_int(method), ptr(out), C.byref(idx))
        assert idx.value == gi[i, 0, 0]
        np.testing.assert_allclose(out[0:2], c[i, 0], rtol=2e-4, atol=1e-4)
        np.testing.assert_allclose(out[4:7], gp[i, 0, 0], rtol=2e-4, atol=1e-3)
        np.testing.assert_allclose(out[7:11], gq[i, 0, 0], rtol=2e-4, atol=2e-3)
        np.testing.assert_allclose(out[2], pe[i, 0, 0], rtol=2e-4, atol=1e-5)
        np.testing.assert_allclose(out[3], re[i, 0, 0], rtol=2e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# ESDF lower-bound pyramid level (cb200_voxel_set.mip): culling must not change any result
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,voxel,seed", [(40, 0.05, 1), (33, 0.06, 2), (17, 0.11, 3)])
def test_voxel_mip_cull_is_exact_on_host(n, voxel, seed):
    """The device sampler compiled for the host: discrete sphere-vs-ESDF cost and gradient with the numpy-built
    lower-bound level are bit-identical to those without it (grid sizes that are and are not multiples of 8, spheres
    inside, at the boundary of and outside the grid), and the level really culls most samples' fetches."""
    from helpers import numpy_voxel_mip
    rng = np.random.default_rng(seed)
    sdf = make_box_esdf(n=n, voxel_size=voxel, num_boxes=6, seed=seed, ground_z=-0.05)
    vox = VoxelWorld.from_grid(sdf.reshape(n, n, n), voxel)
    half = 0.5 * n * voxel
    B, S = 6, 200
    sph = np.zeros((B, 1, S, 4), np.float32)
    sph[..., :3] = rng.uniform(-1.15 * half, 1.15 * half, size=(B, 1, S, 3))
    sph[..., 3] = rng.uniform(0.01, 0.08, size=(B, 1, S))
    sph[0, 0, :5, 3] = -1.0                                   # disabled spheres
    mip = numpy_voxel_mip(vox)
    lib = hostmath()
    c0, g0 = hm_scene(lib, sph, 5000.0, 0.02, False, False, 0.0, vox=vox)
    c1, g1 = hm_scene(lib, sph, 5000.0, 0.02, False, False, 0.0, vox=vox, mip=mip)
    assert np.array_equal(c0, c1) and np.array_equal(g0, g1)
    assert (c0 > 0).any() and (c0 == 0).mean() > 0.3
    want_c, want_g = O.scene_collision(sph, 5000.0, 0.02, None, vox, None, sweep=False, speed_dt=None)
    np.testing.assert_allclose(c1, want_c, rtol=2e-4, atol=1e-5 * want_c.max())
    # the bound is a true lower bound of every trilinear sample based in its block
    feats = vox.features.reshape(-1)[: n * n * n].astype(np.float32).reshape(n, n, n)
    m = mip[0].view(np.float16).astype(np.float32)
    from helpers import mip_block
    Bk = mip_block()
    mm = (n + Bk - 1) // Bk
    for _ in range(200):
        x0, y0, z0 = rng.integers(0, n - 1, 3)
        corner_min = feats[x0:x0 + 2, y0:y0 + 2, z0:z0 + 2].min()
        assert m[((x0 // Bk) * mm + (y0 // Bk)) * mm + (z0 // Bk)] <= corner_min
