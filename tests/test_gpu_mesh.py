"""Mesh obstacles through the C ABI: cb200_sphere_mesh_collision via the drop-in operators of
curobo_b200.scene with a SceneData that carries curobo_b200.mesh.MeshData.

Pins: (1) the reference's own regression for this path -- a box MESH must cost what the analytic CUBOID costs
(tests/_src/collision/test_mesh_collision_sdf.py:17-60) -- here on random spheres, discrete and swept; (2) the oracle's
brute-force restatement of compute_local_sdf_with_grad (oracle/mesh_oracle.py: closest point over ALL triangles + ray-parity
sign, sharing neither the BVH nor the pseudo-normal sign with the product) on an icosphere and on a rotated, translated box,
through the collision cost formula; (3) the analytic sphere."""
import numpy as np
import pytest
import torch

from curobo_b200.mesh import MeshData, MeshWorld, box_mesh, build_bvh, icosphere
from curobo_b200.scene import CollisionBuffer, CuboidData, SceneData, SphereObstacleCollision, SweptSphereObstacleCollision
from curobo_b200.world import CuboidWorld
from oracle import mesh_oracle as MO
from oracle import rollout_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def T(a, dt=None):
    t = torch.as_tensor(np.ascontiguousarray(a)).to(DEV)
    return t.to(dt) if dt is not None else t


def run_scene(sph, scene, w=2.0, eta=0.02, sweep=False, speed_dt=None, env=None, multi=False):
    buf = CollisionBuffer.from_shape(sph.shape, DEV)
    wt, et = T(np.array([w], np.float32)), T(np.array([eta], np.float32))
    if not sweep:
        d = SphereObstacleCollision.apply(T(sph), buf, scene, wt, et, None, env, multi, False)
    else:
        d = SweptSphereObstacleCollision.apply(T(sph), buf, scene, wt, et, None, None if speed_dt is None else T(np.array([speed_dt], np.float32)),
                                               speed_dt is not None, env, multi, False)
    torch.cuda.synchronize() if DEV != "cpu" else None
    return d.cpu().numpy().copy(), buf.gradient.cpu().numpy().copy()


def random_spheres(B, H, S, seed, extent=0.5, r=(0.01, 0.08)):
    rng = np.random.default_rng(seed)
    sph = np.zeros((B, H, S, 4), np.float32)
    base = rng.uniform(-extent, extent, (B, 1, S, 3))
    sph[..., :3] = base + np.cumsum(rng.normal(0, 0.03, (B, H, S, 3)), axis=1)
    sph[..., 3] = rng.uniform(*r, (B, H, S))
    sph[0, 0, 0, 3] = -1.0                      # a disabled sphere
    return sph


@pytest.mark.parametrize("sweep", [False, True])
def test_box_mesh_costs_what_the_cuboid_costs(sweep):
    """The reference's regression (test_mesh_collision_sdf.py): same geometry, mesh vs analytic cuboid -> same cost.  The
    gradients are compared up to the sign the reference gives the mesh gradient outside the surface (data_mesh.py:694-698 returns
    -(cl - p) / |cl - p|, i.e. +d sdf / dp outside, where the cuboid returns -d sdf / dp; reproduced as written)."""
    dims, pose = [0.3, 0.2, 0.25], [0.1, -0.05, 0.02, 0.9238795, 0.0, 0.3826834, 0.0]
    v, f = box_mesh(dims)
    mesh = MeshData.from_world(MeshWorld.create([{"vertices": v, "faces": f, "pose": pose}], max_n=3), DEV)
    cub = CuboidData.from_world(CuboidWorld.create([{"dims": dims, "pose": pose}], max_n=2), DEV)
    sph = random_spheres(5, 6, 40, seed=1)
    kw = dict(sweep=sweep, speed_dt=0.05 if sweep else None)
    dm, gm = run_scene(sph, SceneData(mesh=mesh), **kw)
    dc, gc = run_scene(sph, SceneData(cuboid=cub), **kw)
    assert (dc > 0).sum() > 20 and (dc == 0).sum() > 20
    np.testing.assert_allclose(dm, dc, rtol=2e-4, atol=2e-6 * dc.max())
    assert (dm[0, 0, 0] == 0) and (gm[0, 0, 0] == 0).all()
    if not sweep:      # centres outside the box: opposite gradient sign, same magnitude (inside they agree)
        cen_local = O._quat_rotate(np.broadcast_to(np.array([0, -0.3826834, 0, 0.9238795], np.float32), sph.shape[:-1] + (4,)),
                                   sph[..., :3] - np.array(pose[:3], np.float32))
        outside = (np.abs(cen_local) > 0.5 * np.array(dims, np.float32) + 1e-4).any(-1) & (dc > 0)
        face_region = outside & ((np.abs(cen_local) > 0.5 * np.array(dims, np.float32)).sum(-1) == 1)
        assert face_region.sum() > 5
        np.testing.assert_allclose(gm[face_region][:, :3], -gc[face_region][:, :3], rtol=2e-3, atol=2e-5 * np.abs(gc).max())


def _oracle_discrete_cost(sph, v, f, pose, w, eta):
    """cost / gradient of wp_collision_kernel.py:112-166 with oracle/mesh_oracle.py as the SDF."""
    from curobo_b200.world import _inv_pose_from_pose
    inv = _inv_pose_from_pose(pose)
    ip, iq = O._load_inv_transform(inv)
    c = sph[..., :3].reshape(-1, 3)
    r = sph[..., 3].reshape(-1)
    loc = O._quat_rotate(np.broadcast_to(iq, (c.shape[0], 4)), c) + ip
    cost = np.zeros(c.shape[0], np.float32)
    grad = np.zeros((c.shape[0], 3), np.float32)
    fq = np.array([-iq[0], -iq[1], -iq[2], iq[3]], np.float32)
    for i in range(c.shape[0]):
        if r[i] < 0:
            continue
        radj = np.float32(r[i] + eta)
        sdf, g = MO.mesh_sdf_grad(v, f, loc[i:i + 1], query_distance=float(radj))
        pen = radj - sdf[0]
        if pen > 0:
            ac, ak = O.collision_activation(np.array([pen], np.float32), np.float32(eta))
            cost[i] = w * ac[0]
            grad[i] = w * ak[0] * O._quat_rotate(fq[None], g)[0]
    return cost.reshape(sph.shape[:-1]), grad.reshape(sph.shape[:-1] + (3,))


@pytest.mark.parametrize("shape", ["icosphere", "box"])
def test_mesh_collision_vs_brute_force_oracle(shape):
    if shape == "icosphere":
        v, f = icosphere(0.25, 2)
        pose = [0.05, 0.0, -0.1, 1, 0, 0, 0]
    else:
        v, f = box_mesh([0.4, 0.1, 0.3])
        pose = [0.0, 0.1, 0.0, 0.8660254, 0.0, 0.0, 0.5]
    mesh = MeshData.from_world(MeshWorld.create([{"vertices": v, "faces": f, "pose": pose}]), DEV)
    sph = random_spheres(3, 1, 60, seed=4, extent=0.4)
    d, g = run_scene(sph, SceneData(mesh=mesh), w=3.0, eta=0.03)
    wd, wg = _oracle_discrete_cost(sph, v, f, pose, np.float32(3.0), np.float32(0.03))
    assert (wd > 0).sum() > 15
    np.testing.assert_allclose(d, wd, rtol=5e-4, atol=5e-6 * wd.max())
    np.testing.assert_allclose(g[..., :3], wg, rtol=5e-3, atol=5e-5 * np.abs(wg).max())
    if shape == "icosphere":      # and the analytic sphere: the polyhedron's SDF is within its sagitta of |p| - R
        c = sph[..., :3] - np.array(pose[:3], np.float32)
        sdf_true = np.linalg.norm(c, axis=-1) - 0.25
        inside_deep = (sdf_true < -0.05) & (sph[..., 3] >= 0)
        assert (d[inside_deep] > 0).all()


def test_mesh_with_cuboids_multi_env_and_disabled_slots():
    """Meshes next to cuboids (two launches accumulate), two environments picked by env_query_idx, a disabled mesh slot."""
    v1, f1 = box_mesh([0.2, 0.2, 0.2])
    v2, f2 = icosphere(0.15, 1)
    w = MeshWorld([[{"vertices": v1, "faces": f1, "pose": [0.2, 0, 0, 1, 0, 0, 0]}, {"vertices": v2, "faces": f2, "pose": [-0.2, 0, 0, 1, 0, 0, 0]}],
                   [{"vertices": v2, "faces": f2, "pose": [0.0, 0.2, 0, 1, 0, 0, 0]}]], max_n=2)
    mesh = MeshData.from_world(w, DEV)
    cubw = CuboidWorld.create([{"dims": [0.1, 0.6, 0.1], "pose": [0, 0, 0.2, 1, 0, 0, 0]}], max_n=2)
    cub = CuboidData.from_world(cubw, DEV)
    sph = random_spheres(4, 1, 50, seed=9, extent=0.35)
    env = T(np.array([0, 1, 1, 0], np.int32))
    d_both, g_both = run_scene(sph, SceneData(cuboid=cub, mesh=mesh), env=env, multi=True)
    d_m, g_m = run_scene(sph, SceneData(mesh=mesh), env=env, multi=True)
    d_c, g_c = run_scene(sph, SceneData(cuboid=cub), env=env, multi=True)
    np.testing.assert_allclose(d_both, d_m + d_c, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(g_both, g_m + g_c, rtol=1e-6, atol=1e-7)
    # env 1 only has the sphere mesh at y = 0.2: rows 1, 2 must match a single-mesh world
    single = MeshData.from_world(MeshWorld.create([{"vertices": v2, "faces": f2, "pose": [0.0, 0.2, 0, 1, 0, 0, 0]}]), DEV)
    d_s, _ = run_scene(sph[1:3], SceneData(mesh=single))
    np.testing.assert_array_equal(d_m[1:3], d_s)
    mesh.enable[0, 1] = 0                                   # disable the sphere of env 0
    d_dis, _ = run_scene(sph, SceneData(mesh=mesh), env=env, multi=True)
    only_box = MeshData.from_world(MeshWorld.create([{"vertices": v1, "faces": f1, "pose": [0.2, 0, 0, 1, 0, 0, 0]}]), DEV)
    d_box, _ = run_scene(sph[[0, 3]], SceneData(mesh=only_box))
    np.testing.assert_array_equal(d_dis[[0, 3]], d_box)
