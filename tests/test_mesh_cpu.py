"""CPU tests of the mesh obstacle path: the oracle against analytic shapes, the host BVH builder, and the
product's device routines compiled for the host (tests/hostmath: hm_mesh_sdf = cb200_mesh.cuh::mesh_sdf_grad over a BVH built
by curobo_b200.mesh.build_bvh) against the oracle's brute force."""
import ctypes as C

import numpy as np
import pytest

from curobo_b200.mesh import box_mesh, build_bvh, icosphere
from oracle import mesh_oracle as MO
from oracle import rollout_oracle as O


def test_oracle_on_a_box_equals_the_cuboid_sdf():
    """The reference's regression geometry: the mesh of a box has the SDF of the analytic cuboid (data_cuboid.py:547-628)."""
    dims = [0.3, 0.2, 0.5]
    v, f = box_mesh(dims)
    rng = np.random.default_rng(0)
    p = rng.uniform(-0.3, 0.3, (600, 3)).astype(np.float32)
    sdf, g = MO.mesh_sdf_grad(v, f, p, query_distance=10.0)
    want, wn = O.cuboid_sdf_grad(p, np.array(dims, np.float32))
    np.testing.assert_allclose(sdf, want, rtol=1e-5, atol=1e-6)
    assert (sdf < 0).sum() > 20 and (sdf > 0).sum() > 30
    inside = want < -1e-3
    # inside, both conventions point away from the nearest face; outside the mesh gradient is the negated cuboid gradient
    face_out = (want > 1e-3) & ((np.abs(p) > 0.5 * np.array(dims)).sum(-1) == 1)
    np.testing.assert_allclose(g[face_out], -wn[face_out], atol=1e-5)
    unique_face = inside & (np.sort(np.abs(np.abs(p) - 0.5 * np.array(dims)), -1)[:, 1] - np.sort(np.abs(np.abs(p) - 0.5 * np.array(dims)), -1)[:, 0] > 1e-3)
    np.testing.assert_allclose(g[unique_face], wn[unique_face], atol=1e-5)


def test_oracle_on_an_icosphere_is_close_to_the_sphere():
    v, f = icosphere(0.5, 3)
    rng = np.random.default_rng(1)
    p = rng.uniform(-0.9, 0.9, (300, 3)).astype(np.float32)
    sdf, g = MO.mesh_sdf_grad(v, f, p, query_distance=10.0)
    true = np.linalg.norm(p, axis=-1) - 0.5
    assert np.abs(sdf - true).max() < 0.004                   # sagitta of the level-3 geodesic sphere
    assert ((sdf < 0) == (true < 0))[np.abs(true) > 0.005].all()
    far = MO.mesh_sdf_grad(v, f, np.array([[5.0, 0, 0]], np.float32), query_distance=0.1)
    assert far[0][0] == np.float32(0.5 * np.linalg.norm(v.max(0) - v.min(0))) and (far[1] == 0).all()   # nothing within max_distance


@pytest.fixture(scope="module")
def hm():
    from curobo_b200 import build
    so = build.build_hostmath(force=False, verbose=False)
    L = C.CDLL(so)
    L.hm_mesh_sdf.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
    L.hm_mesh_sdf.restype = None
    return L


@pytest.mark.parametrize("shape", ["box", "icosphere", "tetra"])
def test_device_routine_on_the_host_vs_brute_force(hm, shape):
    """cb200_mesh.cuh::mesh_sdf_grad (stackless BVH traversal, closest point with its feature, pseudo-normal sign) compiled for the
    host, against the oracle (all triangles, ray-parity sign): distances to float32 rounding, signs identical, gradients equal."""
    if shape == "box":
        v, f = box_mesh([0.4, 0.3, 0.2])
    elif shape == "icosphere":
        v, f = icosphere(0.3, 2)
    else:
        v = np.array([[0, 0, 0], [0.3, 0, 0], [0, 0.3, 0], [0, 0, 0.3]], np.float32) - 0.07
        f = np.array([[0, 2, 1], [0, 1, 3], [1, 2, 3], [0, 3, 2]], np.int64)
    nodes, tris = build_bvh(v, f)
    rng = np.random.default_rng(5)
    p = rng.uniform(v.min(0) - 0.12, v.max(0) + 0.12, (800, 3)).astype(np.float32)
    out = np.zeros((p.shape[0], 4), np.float32)
    md = 0.5 * float(np.linalg.norm(v.max(0) - v.min(0)))
    md = max(md, 2.0)
    hm.hm_mesh_sdf(nodes.ctypes.data, tris.ctypes.data, p.ctypes.data, p.shape[0], C.c_float(md), out.ctypes.data)
    sdf, g = MO.mesh_sdf_grad(v, f, p, query_distance=md)
    assert (sdf < 0).sum() > 5
    np.testing.assert_allclose(out[:, 0], sdf, rtol=2e-5, atol=2e-6)
    clear = np.abs(sdf) > 1e-4
    np.testing.assert_allclose(out[clear, 1:], g[clear], atol=2e-4)


def test_bvh_builder_invariants():
    """Host-side structure: depth-first order with skip links, every triangle in exactly one leaf, boxes contain their subtrees."""
    v, f = icosphere(1.0, 3)
    nodes, tris = build_bvh(v, f)
    N, Tn = nodes.shape[0], tris.shape[0]
    assert Tn == f.shape[0] == 1280
    skip = nodes[:, 0, 3].view(np.int32)
    leaf = nodes[:, 1, 3].view(np.int32)
    assert skip[0] == N and (skip > np.arange(N)).all() and (skip <= N).all()
    seen = np.zeros(Tn, int)
    for i in np.nonzero(leaf >= 0)[0]:
        first, cnt = leaf[i] >> 4, leaf[i] & 15
        assert 1 <= cnt <= 4 and skip[i] == i + 1
        seen[first:first + cnt] += 1
        pts = tris[first:first + cnt, :3, :3].reshape(-1, 3)
        assert (pts >= nodes[i, 0, :3]).all() and (pts <= nodes[i, 1, :3]).all()
    assert (seen == 1).all()
    for i in np.nonzero(leaf < 0)[0]:                       # children: i + 1 and skip[i + 1]
        for ch in (i + 1, skip[i + 1]):
            assert ch < skip[i]
            assert (nodes[ch, 0, :3] >= nodes[i, 0, :3]).all() and (nodes[ch, 1, :3] <= nodes[i, 1, :3]).all()
    # same triangles, reordered
    key = lambda t: np.sort(np.round(t[:, :3, :3].reshape(t.shape[0], -1), 6), axis=1)  # noqa: E731
    want = np.stack([v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]], 1).reshape(Tn, -1)
    assert np.allclose(np.sort(key(tris), axis=0), np.sort(np.sort(np.round(want, 6), axis=1), axis=0))
