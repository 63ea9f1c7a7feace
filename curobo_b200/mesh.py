"""Mesh obstacles: host side.

  MeshData     <- curobo/_src/geom/data/data_mesh.py:40-520: per (env, slot) bounding-box `dims [n_env, max_n, 4]`, `inv_pose
                  [n_env, max_n, 8]` (x y z qw qx qy qz pad), `enable`, `count`, `max_n`, `num_envs` -- the same tensors; the
                  reference's `mesh_ids` (handles of Warp meshes, whose BVH lives inside warp-lang) are replaced by this
                  repository's own BVH buffers: `nodes`, `triangles` + per-slot offsets.
  build_bvh    the structure the kernels traverse (curobo_b200/csrc/cb200_mesh.cuh): median-split binary tree, <= 4 triangles per
               leaf, nodes in depth-first order with skip links (stackless traversal), triangles reordered by leaf, every
               triangle carrying its face normal and the angle-weighted pseudo-normals of its edges and vertices (exact
               inside / outside sign for closed manifold meshes, Baerentzen & Aanaes 2005).
The build is numpy on the host (done once per mesh, like the reference's wp.Mesh construction); the queries are CUDA only.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import lib as _lib
from .backends.tensor_checks import check_tensors
from .world import _inv_pose_from_pose

LEAF_SIZE = 4


def _unit(v):
    n = np.linalg.norm(v, axis=-1, keepdims=True)
    return np.where(n > 1e-20, v / np.maximum(n, 1e-20), 0.0)


def triangle_records(vertices: np.ndarray, faces: np.ndarray) -> np.ndarray:
    """[T, 8, 4] float32 triangle records (layout: cb200_mesh.cuh tri_normal): a, b, c, face normal, edge pseudo-normals ab / bc /
    ca in rows 0-6; the vertex pseudo-normals of a, b, c ride in the w lanes of rows 0-2, 3-5 and (6.w, 7.x, 7.y)."""
    V = np.asarray(vertices, np.float64)
    F = np.asarray(faces, np.int64).reshape(-1, 3)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    nf = _unit(np.cross(b - a, c - a))
    # vertex pseudo-normals: face normals weighted by the incident angle
    vn = np.zeros_like(V)
    for i, (p, q, r) in enumerate(((a, b, c), (b, c, a), (c, a, b))):
        u, w = _unit(q - p), _unit(r - p)
        ang = np.arccos(np.clip(np.sum(u * w, -1), -1.0, 1.0))
        np.add.at(vn, F[:, i], nf * ang[:, None])
    # edge pseudo-normals: sum of the normals of the faces sharing the (undirected) edge
    T = F.shape[0]
    ek = np.stack([np.sort(F[:, [0, 1]], 1), np.sort(F[:, [1, 2]], 1), np.sort(F[:, [2, 0]], 1)], 1).reshape(-1, 2)
    key = ek[:, 0] * (V.shape[0] + 1) + ek[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    en_sum = np.zeros((uniq.shape[0], 3))
    np.add.at(en_sum, inv, np.repeat(nf, 3, axis=0))
    en = en_sum[inv].reshape(T, 3, 3)
    rec = np.zeros((T, 8, 4), np.float32)
    rec[:, 0, :3], rec[:, 1, :3], rec[:, 2, :3], rec[:, 3, :3] = a, b, c, nf
    rec[:, 4, :3], rec[:, 5, :3], rec[:, 6, :3] = en[:, 0], en[:, 1], en[:, 2]
    na, nb, nc = vn[F[:, 0]], vn[F[:, 1]], vn[F[:, 2]]
    rec[:, 0, 3], rec[:, 1, 3], rec[:, 2, 3] = na[:, 0], na[:, 1], na[:, 2]
    rec[:, 3, 3], rec[:, 4, 3], rec[:, 5, 3] = nb[:, 0], nb[:, 1], nb[:, 2]
    rec[:, 6, 3], rec[:, 7, 0], rec[:, 7, 1] = nc[:, 0], nc[:, 1], nc[:, 2]
    return rec


def build_bvh(vertices: np.ndarray, faces: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """(nodes [N, 2, 4] float32, triangles [T, 8, 4] float32).  Node i: (box min, skip_i as int bits) (box max, leaf word as int
    bits); its subtree is the index range [i, skip_i); leaf word = -1 for inner nodes, first_triangle * 16 + count for leaves
    (triangles are stored in leaf order).  nodes[0] is the root, whose skip link is N."""
    rec = triangle_records(vertices, faces)
    T = rec.shape[0]
    if T == 0:
        raise ValueError("mesh has no triangles")
    tri = rec[:, :3, :3].astype(np.float64)
    lo, hi, cen = tri.min(1), tri.max(1), tri.mean(1)
    nodes: List[list] = []
    order: List[int] = []

    def rec_build(idx: np.ndarray) -> None:
        me = len(nodes)
        nodes.append(None)
        bmin, bmax = lo[idx].min(0), hi[idx].max(0)
        if idx.shape[0] <= LEAF_SIZE:
            leaf = len(order) * 16 + idx.shape[0]
            order.extend(int(i) for i in idx)
            nodes[me] = [bmin, bmax, leaf, me + 1]
            return
        c = cen[idx]
        ax = int(np.argmax(c.max(0) - c.min(0)))
        srt = idx[np.argsort(c[:, ax], kind="stable")]
        half = srt.shape[0] // 2
        rec_build(srt[:half])
        rec_build(srt[half:])
        nodes[me] = [bmin, bmax, -1, len(nodes)]

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 200))
    try:
        rec_build(np.arange(T))
    finally:
        sys.setrecursionlimit(old)
    N = len(nodes)
    out = np.zeros((N, 2, 4), np.float32)
    skip = np.array([n[3] for n in nodes], np.int32)
    leaf = np.array([n[2] for n in nodes], np.int32)
    # boxes are rounded OUTWARD to float32 so that a triangle never leaves its (float32) box
    bmin = np.stack([n[0] for n in nodes]).astype(np.float64)
    bmax = np.stack([n[1] for n in nodes]).astype(np.float64)
    out[:, 0, :3] = np.nextafter(bmin.astype(np.float32), np.float32(-np.inf))
    out[:, 1, :3] = np.nextafter(bmax.astype(np.float32), np.float32(np.inf))
    out[:, 0, 3] = skip.view(np.float32)
    out[:, 1, 3] = leaf.view(np.float32)
    return out, np.ascontiguousarray(rec[np.asarray(order)])


@dataclass
class MeshWorld:
    """Host description of the mesh obstacles of every environment: per env a list of dicts
    {"vertices": [V,3], "faces": [T,3], "pose": [x y z qw qx qy qz]} (geom/types.py Mesh, triangulated, local frame)."""
    envs: List[List[dict]]
    max_n: int

    @classmethod
    def create(cls, meshes: Sequence[dict], max_n: Optional[int] = None) -> "MeshWorld":
        return cls([list(meshes)], max_n if max_n is not None else max(1, len(meshes)))


@dataclass
class MeshData:
    dims: torch.Tensor          # [n_env, max_n, 4]  bounding-box extents (x, y, z, pad)
    inv_pose: torch.Tensor      # [n_env, max_n, 8]
    enable: torch.Tensor        # [n_env, max_n] uint8
    count: torch.Tensor         # [n_env] int32
    nodes: torch.Tensor         # [sum N, 2, 4] float32
    triangles: torch.Tensor     # [sum T, 8, 4] float32
    node_offset: torch.Tensor   # [n_env * max_n] int32
    triangle_offset: torch.Tensor
    max_n: int
    num_envs: int

    @classmethod
    def from_world(cls, w: MeshWorld, device) -> "MeshData":
        n_env, max_n = len(w.envs), w.max_n
        dims = np.zeros((n_env, max_n, 4), np.float32)
        inv = np.zeros((n_env, max_n, 8), np.float32)
        inv[..., 3] = 1.0
        enable = np.zeros((n_env, max_n), np.uint8)
        count = np.zeros((n_env,), np.int32)
        noff = np.zeros((n_env * max_n,), np.int32)
        toff = np.zeros((n_env * max_n,), np.int32)
        all_nodes, all_tris = [], []
        n_nodes = n_tris = 0
        for e, meshes in enumerate(w.envs):
            if len(meshes) > max_n:
                raise ValueError(f"environment {e} has {len(meshes)} meshes, cache holds {max_n}")
            count[e] = len(meshes)
            for i, m in enumerate(meshes):
                v = np.asarray(m["vertices"], np.float32).reshape(-1, 3)
                f = np.asarray(m["faces"], np.int64).reshape(-1, 3)
                nodes, tris = build_bvh(v, f)
                k = e * max_n + i
                noff[k], toff[k] = n_nodes, n_tris
                n_nodes += nodes.shape[0]
                n_tris += tris.shape[0]
                all_nodes.append(nodes)
                all_tris.append(tris)
                dims[e, i, :3] = v.max(0) - v.min(0)
                inv[e, i] = _inv_pose_from_pose(m.get("pose", (0, 0, 0, 1, 0, 0, 0)))
                enable[e, i] = 1
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(device)  # noqa: E731
        nodes = np.concatenate(all_nodes) if all_nodes else np.zeros((1, 2, 4), np.float32)
        tris = np.concatenate(all_tris) if all_tris else np.zeros((1, 8, 4), np.float32)
        return cls(t(dims), t(inv), t(enable), t(count), t(nodes), t(tris), t(noff), t(toff), max_n, n_env)


def c_mesh_set(d: Optional[MeshData], dev=None) -> Optional[_lib.MeshSet]:
    if d is None:
        return None
    if dev is not None:
        check_tensors(dev, torch.float32, mesh_dims=d.dims, mesh_inv_pose=d.inv_pose, mesh_nodes=d.nodes, mesh_triangles=d.triangles)
        check_tensors(dev, torch.uint8, mesh_enable=d.enable)
        check_tensors(dev, torch.int32, mesh_count=d.count, mesh_node_offset=d.node_offset, mesh_triangle_offset=d.triangle_offset)
    return _lib.MeshSet(d.nodes.data_ptr(), d.triangles.data_ptr(), d.node_offset.data_ptr(), d.triangle_offset.data_ptr(),
                        d.dims.data_ptr(), d.inv_pose.data_ptr(), d.enable.data_ptr(), d.count.data_ptr(), int(d.max_n),
                        int(d.num_envs))


# ------------------------------------------------------------------------------------------------ test / bench geometry
def box_mesh(dims: Sequence[float]) -> Tuple[np.ndarray, np.ndarray]:
    """Closed, outward-oriented triangle mesh of an axis-aligned box centred at the origin (12 triangles)."""
    hx, hy, hz = (0.5 * float(d) for d in dims)
    v = np.array([[-hx, -hy, -hz], [hx, -hy, -hz], [hx, hy, -hz], [-hx, hy, -hz],
                  [-hx, -hy, hz], [hx, -hy, hz], [hx, hy, hz], [-hx, hy, hz]], np.float32)
    f = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4],
                  [1, 2, 6], [1, 6, 5], [2, 3, 7], [2, 7, 6], [3, 0, 4], [3, 4, 7]], np.int64)
    return v, f


def icosphere(radius: float = 1.0, subdivisions: int = 2) -> Tuple[np.ndarray, np.ndarray]:
    """Closed, outward-oriented geodesic sphere (20 * 4^subdivisions triangles)."""
    t = (1.0 + 5.0 ** 0.5) / 2.0
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
         [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = [np.asarray(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(i, j):
            k = (min(i, j), max(i, j))
            if k not in cache:
                m = v[i] + v[j]
                v.append(m / np.linalg.norm(m))
                cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    return (np.asarray(v) * radius).astype(np.float32), np.asarray(f, np.int64)
