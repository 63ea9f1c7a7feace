"""TEST INFRASTRUCTURE -- CPU restatement of the mesh obstacle SDF.

Only tests/ may import this module.  The reference computes it with Warp's mesh queries -- compute_local_sdf_with_grad,
curobo/_src/geom/data/data_mesh.py:643-700:
    max_distance = max(|bounding box| / 2, query_distance)
    result = wp.mesh_query_point(mesh, p, max_distance)            # closest point within max_distance + inside / outside sign
    not found        -> (max_distance, 0, 0, 0)
    cl = closest point;  delta = cl - p;  signed_dist = |delta| * sign;  grad_local = -delta / |delta|  (0 below 1e-6)
warp-lang is a third-party dependency (pyproject.toml:37, unpinned >= 0.10, not vendored, not installable here): its
mesh_query_point is documented as "closest point on the mesh + sign (negative inside)" and is restated from that contract: brute
force over ALL triangles (exact closest point on a triangle, Ericson, Real-Time Collision Detection 5.1.5) and the sign from an
independent inside test -- ray-crossing parity along three axes with a majority vote -- so that the product's BVH traversal and its
pseudo-normal sign are checked against something that shares neither.  Parity at the Warp boundary is unpinned (no Warp here); the
reference's own regression for this path -- a box mesh must cost what the analytic cuboid costs
(tests/_src/collision/test_mesh_collision_sdf.py:17-60) -- is restated in tests/test_mesh_cpu.py and tests/test_gpu_mesh.py.
float64 arithmetic, results cast to float32.
"""
import numpy as np


def closest_points_on_triangles(p, a, b, c):
    """p [N,3]; a, b, c [T,3] -> closest point [N,T,3] on every triangle."""
    p = p[:, None, :]
    a, b, c = a[None], b[None], c[None]
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = np.sum(ab * ap, -1), np.sum(ac * ap, -1)
    bp = p - b
    d3, d4 = np.sum(ab * bp, -1), np.sum(ac * bp, -1)
    cp = p - c
    d5, d6 = np.sum(ab * cp, -1), np.sum(ac * cp, -1)
    vc = d1 * d4 - d3 * d2
    vb = d5 * d2 - d1 * d6
    va = d3 * d6 - d5 * d4
    with np.errstate(divide="ignore", invalid="ignore"):
        denom = 1.0 / (va + vb + vc)
        face = a + ab * (vb * denom)[..., None] + ac * (vc * denom)[..., None]
        e_ab = a + ab * (d1 / (d1 - d3))[..., None]
        e_ac = a + ac * (d2 / (d2 - d6))[..., None]
        e_bc = b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)))[..., None]
    out = face
    conds = [((va <= 0) & ((d4 - d3) >= 0) & ((d5 - d6) >= 0), e_bc), ((vb <= 0) & (d2 >= 0) & (d6 <= 0), e_ac),
             ((d6 >= 0) & (d5 <= d6), np.broadcast_to(c, face.shape)), ((vc <= 0) & (d1 >= 0) & (d3 <= 0), e_ab),
             ((d3 >= 0) & (d4 <= d3), np.broadcast_to(b, face.shape)), ((d1 <= 0) & (d2 <= 0), np.broadcast_to(a, face.shape))]
    for cond, val in conds:          # later entries win: same precedence as the sequential tests of the textbook routine
        out = np.where(cond[..., None], val, out)
    return out


def inside_by_ray_parity(p, a, b, c):
    """Point-in-closed-mesh by crossing parity of axis-aligned rays (+x, +y, +z), majority vote."""
    votes = np.zeros(p.shape[0], np.int64)
    for ax in range(3):
        u, v = (ax + 1) % 3, (ax + 2) % 3
        pu, pv, pa = p[:, None, u], p[:, None, v], p[:, None, ax]
        au, av, bu, bv, cu, cv = a[None, :, u], a[None, :, v], b[None, :, u], b[None, :, v], c[None, :, u], c[None, :, v]
        den = (bv - cv) * (au - cu) + (cu - bu) * (av - cv)
        with np.errstate(divide="ignore", invalid="ignore"):
            l1 = ((bv - cv) * (pu - cu) + (cu - bu) * (pv - cv)) / den
            l2 = ((cv - av) * (pu - cu) + (au - cu) * (pv - cv)) / den
            l3 = 1.0 - l1 - l2
            hit_a = l1 * a[None, :, ax] + l2 * b[None, :, ax] + l3 * c[None, :, ax]
        hit = (np.abs(den) > 1e-300) & (l1 >= 0) & (l2 >= 0) & (l3 >= 0) & (hit_a > pa)
        votes += (hit.sum(1) % 2 == 1)
    return votes >= 2


def mesh_sdf_grad(vertices, faces, p, query_distance=0.0):
    """(sdf [N], grad_local [N,3]) of data_mesh.py:643-700 for points p [N,3] in the mesh's frame."""
    V = np.asarray(vertices, np.float64)
    F = np.asarray(faces, np.int64).reshape(-1, 3)
    p = np.asarray(p, np.float64).reshape(-1, 3)
    a, b, c = V[F[:, 0]], V[F[:, 1]], V[F[:, 2]]
    max_distance = max(0.5 * float(np.linalg.norm(V.max(0) - V.min(0))), float(query_distance))
    cl = closest_points_on_triangles(p, a, b, c)
    d2 = np.sum((cl - p[:, None, :]) ** 2, -1)
    k = np.argmin(d2, 1)
    best = cl[np.arange(p.shape[0]), k]
    delta = best - p
    dist = np.sqrt(np.sum(delta * delta, -1))
    sign = np.where(inside_by_ray_parity(p, a, b, c), -1.0, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        grad = np.where((dist > 1e-6)[:, None], -delta / dist[:, None], 0.0)
    found = dist < max_distance
    sdf = np.where(found, dist * sign, max_distance)
    grad = np.where(found[:, None], grad, 0.0)
    return sdf.astype(np.float32), grad.astype(np.float32)
