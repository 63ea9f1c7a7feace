"""CPU ORACLE for the rollout cost+gradient hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (curobo_b200/) never imports it and has no CPU fallback.

It is a float32 numpy restatement of the reference's arithmetic; every function cites the reference
file:line it follows.  It is vectorised over evaluations but keeps the reference's formulas
(including the places where the reference's "gradient" is hand-defined and not the true derivative).

Pinning status (see DESIGN.md "Oracle"):
  * FK:            pinned by the reference's Franka golden vector (tests/_src/robot/kinematics/
                   test_kinematics.py:57-82) -> tests/test_oracle_golden.py, and on the GPU box against the
                   reference's own CUDA kernels compiled from /root/reference into oracle/_ref.
  * FK backward,   pinned on the GPU box against oracle/_ref (reference kinematics_backward_kernel,
    self-collision self_collision_max_distance_kernel) and by finite differences.
  * scene collision / tool pose / c-space: the reference implements these in NVIDIA Warp (third-party,
    `warp-lang>=0.10`, pyproject.toml:37, not vendored, not installable here), so its kernels cannot be
    compiled or run.  They are pinned on the reference's own SOURCE instead: tests/golden/make_warp_golden.py
    imports the reference's @wp.kernel / @wp.func Python bodies and executes them on the CPU, one thread at
    a time, under a pure-Python stand-in for Warp's builtins (oracle/warp_shim); the resulting fixture
    (tests/golden/warp_reference_golden.npz: discrete / multi-env / swept collision, speed metric, tool pose
    axis-angle and Lie, c-space STATE with a live effort channel, c-space POSITION) is reproduced by this
    module to float32 rounding (tests/test_warp_reference_golden_cpu.py).  What remains restated is the
    stand-in's ~35 builtins (quaternion / transform algebra per Warp's documentation, C integer division);
    the reference's property tests (tests/_src/geom/sdf/test_voxel_collision.py:636-1156) are restated in
    tests/ as well.
"""
from __future__ import annotations

import numpy as np

F = np.float32
FIXED, X_PRISM, Y_PRISM, Z_PRISM, X_ROT, Y_ROT, Z_ROT = -1, 0, 1, 2, 3, 4, 5


# ==========================================================================================
# Forward kinematics
# ==========================================================================================

def local_link_transforms(rm, q):
    """local[l] = fixed[l] * J(theta_l), joint motion applied on the right.

    kinematics_forward_helper.cuh:316-393 (compute_local_link_transform);
    theta = offset.x * q[joint_map[l]] + offset.y  (kinematics_util.cuh:62-75).
    q: [N,D] -> [N,nl,3,4]"""
    q = np.asarray(q, F)
    N = q.shape[0]
    nl = rm.link_map.shape[0]
    out = np.broadcast_to(rm.fixed_transforms.astype(F), (N, nl, 3, 4)).copy()
    for l in range(nl):
        jt = int(rm.joint_map_type[l])
        if jt == FIXED:
            continue
        f = rm.fixed_transforms[l].astype(F)
        ang = (F(rm.joint_offset_map[l, 0]) * q[:, int(rm.joint_map[l])] + F(rm.joint_offset_map[l, 1])).astype(F)
        if jt <= Z_PRISM:
            out[:, l, :, 3] = f[:, 3][None, :] + f[:, jt][None, :] * ang[:, None]
            continue
        s, c = np.sin(ang).astype(F), np.cos(ang).astype(F)
        a = jt - X_ROT
        i, j = (a + 1) % 3, (a + 2) % 3          # rotate the two other columns
        # col_i' = c*col_i + s*col_j ; col_j' = c*col_j - s*col_i   (:379-391)
        out[:, l, :, i] = f[:, i][None, :] * c[:, None] + f[:, j][None, :] * s[:, None]
        out[:, l, :, j] = f[:, j][None, :] * c[:, None] - f[:, i][None, :] * s[:, None]
    return out


def compose_chain(rm, local):
    """cumul[0] = fixed[0]; cumul[l] = cumul[link_map[l]] * local[l], l = 1..nl-1 in index order.

    kinematics_forward_helper.cuh:437-512."""
    N, nl = local.shape[:2]
    cum = np.zeros((N, nl, 3, 4), F)
    cum[:, 0] = rm.fixed_transforms[0].astype(F)
    for l in range(1, nl):
        P = cum[:, int(rm.link_map[l])]
        L = local[:, l]
        R = np.einsum("nij,njk->nik", P[:, :, :3], L[:, :, :3]).astype(F)
        t = (np.einsum("nij,nj->ni", P[:, :, :3], L[:, :, 3]) + P[:, :, 3]).astype(F)
        cum[:, l, :, :3] = R
        cum[:, l, :, 3] = t
    return cum


def quat_from_rotation(R):
    """Rotation matrix [..,3,3] -> quaternion wxyz with w >= 0.

    quaternion_util.cuh:110-158 (branchy trace method, negated then sign-normalised :52-58)."""
    R = np.asarray(R, F)
    t = R.reshape(R.shape[:-2] + (9,))
    t0, t1, t2, t3, t4, t5, t6, t7, t8 = [t[..., i] for i in range(9)]
    q = np.zeros(R.shape[:-2] + (4,), F)  # x,y,z,w
    c_a = (t8 < 0) & (t0 > t4)
    c_b = (t8 < 0) & ~(t0 > t4)
    c_c = ~(t8 < 0) & (t0 < -t4)
    c_d = ~(t8 < 0) & ~(t0 < -t4)
    with np.errstate(invalid="ignore", divide="ignore"):
        for cond, n, comps in (
            (c_a, 1 + t0 - t4 - t8, lambda n, s: (n * s, (t1 + t3) * s, (t6 + t2) * s, -(t5 - t7) * s)),
            (c_b, 1 - t0 + t4 - t8, lambda n, s: ((t1 + t3) * s, n * s, (t5 + t7) * s, -(t6 - t2) * s)),
            (c_c, 1 - t0 - t4 + t8, lambda n, s: ((t6 + t2) * s, (t5 + t7) * s, n * s, -(t1 - t3) * s)),
            (c_d, 1 + t0 + t4 + t8, lambda n, s: ((t5 - t7) * s, (t6 - t2) * s, (t1 - t3) * s, -n * s)),
        ):
            s = (F(0.5) / np.sqrt(n.astype(F))).astype(F)
            x, y, z, w = comps(n.astype(F), s)
            for k, v in enumerate((x, y, z, w)):
                q[..., k] = np.where(cond, v, q[..., k])
    inv = (F(1.0) / np.sqrt(np.sum(q * q, axis=-1))).astype(F)
    inv = np.where(q[..., 3] < 0, -inv, inv)
    q = (q * inv[..., None]).astype(F)
    return np.stack([q[..., 3], q[..., 0], q[..., 1], q[..., 2]], axis=-1)  # wxyz


def fk_forward(rm, q, env_query_idx=None, horizon=1):
    """FK + robot spheres + tool poses.  kinematics_forward_kernel.cuh:131-261.

    q [N,D] -> cumul [N,nl,3,4], spheres [N,S,4], link_pos [N,L,3], link_quat [N,L,4] (wxyz, w>=0)
    Sphere p' = R p + t, radius copied (kinematics_util.cuh:39-49).  With several sphere configs
    (link_spheres [n_cfg,S,4]) row n uses config env_query_idx[n // horizon]
    (kinematics_forward_helper.cuh:232-233)."""
    q = np.asarray(q, F)
    cum = compose_chain(rm, local_link_transforms(rm, q))
    ls = np.asarray(rm.link_spheres, F)
    N = q.shape[0]
    if ls.ndim == 3:
        if ls.shape[0] > 1:
            cfg = np.asarray(env_query_idx)[np.arange(N) // horizon]
            ls = ls[cfg]                         # [N,S,4]
        else:
            ls = ls[0][None]
    else:
        ls = ls[None]
    T = cum[:, rm.link_sphere_idx_map.astype(np.int64)]      # [N,S,3,4]
    p = (np.einsum("nsij,nsj->nsi", T[..., :3], np.broadcast_to(ls[..., :3], T.shape[:2] + (3,))) + T[..., 3]).astype(F)
    sph = np.concatenate([p, np.broadcast_to(ls[..., 3:4], p.shape[:2] + (1,))], axis=-1).astype(F)
    Tt = cum[:, rm.tool_frame_map.astype(np.int64)]
    return cum, sph, Tt[..., 3].copy(), quat_from_rotation(Tt[..., :3])


# ==========================================================================================
# FK backward (J^T g)
# ==========================================================================================

def quat_grad_to_omega(quat_wxyz, g_wxyz):
    """omega = 1/2 E(q)^T g.  quaternion_util.cuh:86-103 (quat held xyzw there; same numbers)."""
    w, x, y, z = [quat_wxyz[..., i] for i in range(4)]
    gw, gx, gy, gz = [g_wxyz[..., i] for i in range(4)]
    return np.stack([F(0.5) * (-x * gw + w * gx + z * gy - y * gz),
                     F(0.5) * (-y * gw - z * gx + w * gy + x * gz),
                     F(0.5) * (-z * gw + y * gx - x * gy + w * gz)], axis=-1).astype(F)


def fk_backward(rm, cum, grad_spheres, grad_link_pos, grad_link_quat, env_query_idx=None, horizon=1):
    """grad_q[N,D] = sum_spheres J^T g + sum_tool_frames J^T (g_pos, omega(g_quat)).

    kinematics_backward_kernel.cuh:34-160; kinematics_backward_helper.cuh:15-183;
    kinematics_joint_util.cuh:12-67:
      revolute  joint at link j (axis a = column type-3 of cumul[j], origin o_j, sign s=offset.x):
          g_q += s * g_p . (a x (p - o_j))   (+ s * a . omega for tool frames)
      prismatic: g_q += s * a . g_p
    The sphere position is re-derived from cumul and the link-frame sphere (:60-62); the radius
    component of grad_spheres is ignored (:45-55)."""
    N = cum.shape[0]
    D = len(rm.joint_names) if hasattr(rm, "joint_names") else int(rm.joint_map.max()) + 1
    gq = np.zeros((N, D), F)
    off, co = rm.link_chain_offsets.astype(np.int64), rm.link_chain_data.astype(np.int64)

    def accumulate(link, p, g, omega=None):
        # p, g: [N,K,3] points rigidly attached to `link` with gradient g
        for j in co[off[link]:off[link + 1]]:
            jt = int(rm.joint_map_type[j])
            if jt == FIXED:
                continue
            s = F(rm.joint_offset_map[j, 0])
            d = int(rm.joint_map[j])
            if jt >= X_ROT:
                a = cum[:, j, :, jt - X_ROT]                      # [N,3]
                o = cum[:, j, :, 3]
                r = p - o[:, None, :]
                c = np.cross(np.broadcast_to(a[:, None, :], r.shape), r).astype(F)
                contrib = np.sum((s * g) * c, axis=-1)
                if omega is not None:
                    contrib = contrib + s * np.sum(a[:, None, :] * omega, axis=-1)
            else:
                a = cum[:, j, :, jt]
                contrib = s * np.sum(a[:, None, :] * g, axis=-1)
            gq[:, d] += np.sum(contrib, axis=1).astype(F)

    if grad_spheres is not None:
        ls = np.asarray(rm.link_spheres, F)
        if ls.ndim == 3:
            if ls.shape[0] > 1:
                ls = ls[np.asarray(env_query_idx)[np.arange(N) // horizon]]
            else:
                ls = ls[0][None]
        else:
            ls = ls[None]
        smap = rm.link_sphere_idx_map.astype(np.int64)
        for link in np.unique(smap):
            ids = np.nonzero(smap == link)[0]
            T = cum[:, link]
            lp = np.broadcast_to(ls[:, ids, :3], (N, len(ids), 3))
            p = (np.einsum("nij,nkj->nki", T[:, :, :3], lp) + T[:, None, :, 3]).astype(F)
            accumulate(int(link), p, np.asarray(grad_spheres, F)[:, ids, :3])
    if grad_link_pos is not None:
        for e, link in enumerate(rm.tool_frame_map.astype(np.int64)):
            T = cum[:, link]
            quat = quat_from_rotation(T[:, :, :3])
            om = quat_grad_to_omega(quat, np.asarray(grad_link_quat, F)[:, e])
            accumulate(int(link), T[:, None, :, 3], np.asarray(grad_link_pos, F)[:, e][:, None, :], om[:, None, :])
    return gq


# ==========================================================================================
# Self collision
# ==========================================================================================

def self_collision(spheres, sphere_padding, pairs, weight, chunk=256):
    """Worst-pair self-collision cost and its (hand-defined) gradient.

    self_collision_helper.cuh:61-71 (f = (ri+rj)^2 - |pi-pj|^2, radii padded :180-183),
    :227-277 (running max starts at {0,0,0}; pair valid only if both padded radii >= 0),
    :280-349 (cost = 0.5*w*f_max if f_max>0 else 0; grad_i = w*(p_j-p_i), .w = -w; grad_j = -grad_i,
    .w = -w; every other row zero).  Ties: first pair in list order (collision_pair.cuh:55-57 keeps the
    incumbent on equality).
    spheres [N,S,4] -> cost [N], grad [N,S,4], worst pair index [N] (-1 if none)."""
    sph = np.asarray(spheres, F)
    N, S, _ = sph.shape
    pairs = np.asarray(pairs).astype(np.int64)
    w = F(weight)
    cost = np.zeros(N, F)
    grad = np.zeros((N, S, 4), F)
    best = np.full(N, -1, np.int64)
    pad = np.asarray(sphere_padding, F)
    for n0 in range(0, N, chunk):
        s = sph[n0:n0 + chunk]
        r = (s[..., 3] + pad[None, :]).astype(F)
        pi, pj = s[:, pairs[:, 0], :3], s[:, pairs[:, 1], :3]
        ri, rj = r[:, pairs[:, 0]], r[:, pairs[:, 1]]
        d = (pi - pj).astype(F)
        rs = (ri + rj).astype(F)
        f = (rs * rs - np.sum(d * d, axis=-1)).astype(F)
        f = np.where((ri >= 0) & (rj >= 0), f, F(0.0))
        k = np.argmax(f, axis=1)
        fm = f[np.arange(f.shape[0]), k]
        hit = fm > 0
        idx = np.nonzero(hit)[0]
        cost[n0 + idx] = F(0.5) * w * fm[idx]
        best[n0 + idx] = k[idx]
        i, j = pairs[k[idx], 0], pairs[k[idx], 1]
        g = (w * (s[idx, j, :3] - s[idx, i, :3])).astype(F)
        grad[n0 + idx, i, :3] = g
        grad[n0 + idx, i, 3] = -w
        grad[n0 + idx, j, :3] = -g
        grad[n0 + idx, j, 3] = -w
    return cost, grad, best


# ==========================================================================================
# Scene collision: rigid transforms (Warp semantics), SDFs, activation
# ==========================================================================================

def _quat_rotate(qxyzw, v):
    """wp.quat_rotate: v*(2w^2-1) + 2w(qv x v) + 2 qv (qv.v)   (warp builtin, quat stored x,y,z,w)."""
    qv, w = qxyzw[..., :3], qxyzw[..., 3:4]
    return (v * (F(2.0) * w * w - F(1.0)) + np.cross(qv, v) * w * F(2.0)
            + qv * np.sum(qv * v, axis=-1, keepdims=True) * F(2.0)).astype(F)


def _load_inv_transform(inv_pose_row):
    """inv_pose = [x,y,z,qw,qx,qy,qz,pad] -> (p, q_xyzw).  geom/data/helper_pose.py:36-91."""
    ip = np.asarray(inv_pose_row, F)
    return ip[:3], np.array([ip[4], ip[5], ip[6], ip[3]], F)


def collision_activation(pen, eta):
    """(cost, slope): pen<=0 -> 0; pen>eta -> (pen-eta/2, 1); else (pen^2/(2 eta), pen/eta).

    geom/collision/wp_collision_common.py:12-37.  eta = 0 is legal (always the linear branch)."""
    pen = np.asarray(pen, F)
    eta = F(eta)
    with np.errstate(divide="ignore", invalid="ignore"):
        quad_c = (F(0.5) * pen * pen / eta).astype(F)
        quad_s = (pen / eta).astype(F)
    lin = pen > eta
    c = np.where(lin, pen - F(0.5) * eta, quad_c)
    s = np.where(lin, F(1.0), quad_s)
    pos = pen > 0
    return np.where(pos, c, F(0)).astype(F), np.where(pos, s, F(0)).astype(F)


def cuboid_sdf_grad(p, dims):
    """Box SDF + 'gradient' n = -d sdf/dp (outside) / inward face normal rule (inside).

    geom/data/data_cuboid.py:547-628.  p [...,3] local point, dims (3,) full extents."""
    h = (np.asarray(dims, F)[:3] * F(0.5)).astype(F)
    qv = (np.abs(p) - h).astype(F)
    c = np.maximum(qv, F(0))
    od = np.sqrt(np.sum(c * c, axis=-1)).astype(F)
    maxq = np.max(qv, axis=-1)
    sdf = (od + np.minimum(maxq, F(0))).astype(F)
    eps = F(1e-6)
    neg = p < 0
    with np.errstate(divide="ignore", invalid="ignore"):
        g_out = (c * (F(-1.0) / od)[..., None]).astype(F)
    g_out = np.where(neg, -g_out, g_out)
    # inside: axis with largest q, x before y before z within eps
    is_x = np.abs(qv[..., 0] - maxq) < eps
    is_y = ~is_x & (np.abs(qv[..., 1] - maxq) < eps)
    is_z = ~is_x & ~is_y
    sel = np.stack([is_x, is_y, is_z], axis=-1)
    g_in = np.where(sel, np.where(neg, F(1.0), F(-1.0)), F(0.0)).astype(F)
    g = np.where((od > eps)[..., None], g_out, g_in).astype(F)
    return sdf, g


def voxel_sdf_grad(p, feat, nx, ny, nz, vs, max_dist):
    """Trilinear ESDF value + normalised negative gradient.

    geom/data/data_voxel.py:790-1069 (sample_voxel_sdf_with_grad: v = p/vs + dims/2 - 0.5, floor,
    8 fp16 corners, all-valid fast path :871-917, validity-weighted boundary path :919-1069) and
    :1163-1215 (sdf >= max_dist -> no collision; n = -grad/|grad|, zero if |grad| <= 1e-6).
    p [...,3] f32 local; feat flat fp16 [nx*ny*nz] C-order (z fastest)."""
    p = np.asarray(p, F)
    shp = p.shape[:-1]
    p = p.reshape(-1, 3)
    md = F(max_dist)
    inv = (F(1.0) / F(vs)).astype(F)
    v = [(p[:, a] * inv + F(n) * F(0.5) - F(0.5)).astype(F) for a, n in enumerate((nx, ny, nz))]
    i0 = [np.floor(x).astype(np.int64) for x in v]
    fr = [(x - i.astype(F)).astype(F) for x, i in zip(v, i0)]
    fr1 = [(F(1.0) - f).astype(F) for f in fr]
    dims = (nx, ny, nz)
    ok0 = [(i >= 0) & (i < n) for i, n in zip(i0, dims)]
    ok1 = [((i + 1) >= 0) & ((i + 1) < n) for i, n in zip(i0, dims)]
    sx, sy = ny * nz, nz
    s, val = {}, {}
    featf = np.asarray(feat).reshape(-1)
    for cx in (0, 1):
        for cy in (0, 1):
            for cz in (0, 1):
                ok = (ok1[0] if cx else ok0[0]) & (ok1[1] if cy else ok0[1]) & (ok1[2] if cz else ok0[2])
                idx = (i0[0] + cx) * sx + (i0[1] + cy) * sy + (i0[2] + cz)
                vals = featf[np.where(ok, idx, 0)].astype(F)
                s[(cx, cy, cz)] = np.where(ok, vals, md).astype(F)
                val[(cx, cy, cz)] = ok
    fx, fy, fz = fr
    fx1, fy1, fz1 = fr1
    wgt = {(cx, cy, cz): ((fx if cx else fx1) * (fy if cy else fy1) * (fz if cz else fz1)).astype(F)
           for cx in (0, 1) for cy in (0, 1) for cz in (0, 1)}
    order = [(0, 0, 0), (0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 0), (1, 0, 1), (1, 1, 0), (1, 1, 1)]
    all_valid = np.ones(p.shape[0], bool)
    for k in order:
        all_valid &= val[k]
    # fast path
    sdf_f = np.zeros(p.shape[0], F)
    for k in order:
        sdf_f = (sdf_f + s[k] * wgt[k]).astype(F)

    def bil(pairs_w):
        acc = np.zeros(p.shape[0], F)
        for (hi, lo, w) in pairs_w:
            acc = (acc + (s[hi] - s[lo]) * w).astype(F)
        return acc
    gx_pairs = [((1, 0, 0), (0, 0, 0), fy1 * fz1), ((1, 0, 1), (0, 0, 1), fy1 * fz),
                ((1, 1, 0), (0, 1, 0), fy * fz1), ((1, 1, 1), (0, 1, 1), fy * fz)]
    gy_pairs = [((0, 1, 0), (0, 0, 0), fx1 * fz1), ((0, 1, 1), (0, 0, 1), fx1 * fz),
                ((1, 1, 0), (1, 0, 0), fx * fz1), ((1, 1, 1), (1, 0, 1), fx * fz)]
    gz_pairs = [((0, 0, 1), (0, 0, 0), fx1 * fy1), ((0, 1, 1), (0, 1, 0), fx1 * fy),
                ((1, 0, 1), (1, 0, 0), fx * fy1), ((1, 1, 1), (1, 1, 0), fx * fy)]
    g_f = [(bil(pp) * inv).astype(F) for pp in (gx_pairs, gy_pairs, gz_pairs)]
    # boundary path
    wsum = np.zeros(p.shape[0], F)
    vsum = np.zeros(p.shape[0], F)
    for k in order:
        m = val[k].astype(F)
        vsum = (vsum + s[k] * wgt[k] * m).astype(F)
        wsum = (wsum + wgt[k] * m).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        sdf_b = (vsum / wsum).astype(F)

        def bil_b(pairs_w):
            gs = np.zeros(p.shape[0], F)
            gw = np.zeros(p.shape[0], F)
            for (hi, lo, w) in pairs_w:
                m = val[hi] & val[lo]
                gs = np.where(m, gs + (s[hi] - s[lo]) * w, gs).astype(F)
                gw = np.where(m, gw + w, gw).astype(F)
            return np.where(gw > 0, gs / gw * inv, F(0)).astype(F)
        g_b = [bil_b(pp) for pp in (gx_pairs, gy_pairs, gz_pairs)]
    none_valid = wsum <= 0
    sdf = np.where(all_valid, sdf_f, np.where(none_valid, md, sdf_b)).astype(F)
    g = np.stack([np.where(all_valid, gf, np.where(none_valid, F(0), gb)) for gf, gb in zip(g_f, g_b)], axis=-1).astype(F)
    # compute_local_sdf_with_grad :1163-1215
    far = sdf >= md
    n = -g
    ln = np.sqrt(np.sum(n * n, axis=-1)).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        nn = np.where((ln > F(1e-6))[:, None], n / ln[:, None], F(0)).astype(F)
    sdf = np.where(far, md, sdf).astype(F)
    nn = np.where(far[:, None], F(0), nn).astype(F)
    return sdf.reshape(shp), nn.reshape(shp + (3,))


def _obstacles(world_cuboid, world_voxel, env):
    """Yield (kind, inv_pose_row, sdf_fn) for every enabled obstacle of environment `env`.

    is_obs_enabled: local_idx < count[env] and enable == 1  (data_cuboid.py / data_voxel.py:1083-1101)."""
    if world_cuboid is not None:
        for i in range(min(int(world_cuboid.count[env]), world_cuboid.dims.shape[1])):
            if world_cuboid.enable[env, i] != 1:
                continue
            dims = world_cuboid.dims[env, i]
            yield world_cuboid.inv_pose[env, i], (lambda p, dims=dims: cuboid_sdf_grad(p, dims))
    if world_voxel is not None:
        for i in range(min(int(world_voxel.count[env]), world_voxel.params.shape[1])):
            if world_voxel.enable[env, i] != 1:
                continue
            pr = world_voxel.params[env, i]
            nx, ny, nz, vs = int(pr[0]), int(pr[1]), int(pr[2]), F(pr[3])
            feat = world_voxel.features[env, i]
            md = world_voxel.max_dist
            yield world_voxel.inv_pose[env, i], (lambda p, a=(feat, nx, ny, nz, vs, md): voxel_sdf_grad(p, *a))


def scene_collision(spheres, weight, eta, world_cuboid=None, world_voxel=None, env_query_idx=None,
                    sweep=False, speed_dt=None, sweep_steps=3, stats=None):
    """Per-sphere scene-collision cost [B,H,S] and gradient [B,H,S,4] summed over obstacles.

    discrete: geom/collision/wp_collision_kernel.py:70-166
    swept:    geom/collision/wp_sweep_collision_kernel.py:83-260 (<=3 adaptive samples toward h-1 and h+1,
              accumulated in the obstacle frame, rotated once)
    speed metric (if speed_dt is not None): geom/collision/wp_speed_metric.py:10-93
    Spheres with r < 0 are skipped.  pen = (r + eta) - sdf.
    `stats` (dict, optional): accumulates "samples" = SDF evaluations an adaptive implementation performs (1 at the sphere +
    one per live sweep step) and "sphere_obstacle_pairs", for the measured mean n_s of SURVEY.md 8d."""
    sph = np.asarray(spheres, F)
    B, H, S, _ = sph.shape
    w, eta = F(weight), F(eta)
    cost = np.zeros((B, H, S), F)
    grad = np.zeros((B, H, S, 4), F)
    envs = np.zeros(B, np.int64) if env_query_idx is None else np.asarray(env_query_idx).astype(np.int64)
    for env in np.unique(envs):
        bsel = np.nonzero(envs == env)[0]
        sp = sph[bsel]
        active = sp[..., 3] >= 0
        radj = (sp[..., 3] + eta).astype(F)
        for inv_row, sdf_fn in _obstacles(world_cuboid, world_voxel, int(env)):
            ip, iq = _load_inv_transform(inv_row)
            iq_b = np.broadcast_to(iq, sp.shape[:-1] + (4,))
            loc = (_quat_rotate(iq_b, sp[..., :3]) + ip).astype(F)        # wp.transform_point
            fq = np.array([-iq[0], -iq[1], -iq[2], iq[3]], F)             # transform_inverse rotation
            fq_b = np.broadcast_to(fq, sp.shape[:-1] + (4,))
            sdf, n = sdf_fn(loc)
            pen = (radj - sdf).astype(F)
            c, k = collision_activation(pen, eta)
            if stats is not None:
                stats["samples"] = stats.get("samples", 0) + int(active.sum())
                stats["sphere_obstacle_pairs"] = stats.get("sphere_obstacle_pairs", 0) + int(active.sum())
            if not sweep:
                hit = active & (pen > 0)
                gw = _quat_rotate(fq_b, n)
                cost[bsel] += np.where(hit, w * c, F(0)).astype(F)
                grad[bsel, ..., :3] += np.where(hit[..., None], (w * k)[..., None] * gw, F(0)).astype(F)
                continue
            csum = np.where(pen > 0, c, F(0)).astype(F)
            gsum = np.where((pen > 0)[..., None], k[..., None] * n, F(0)).astype(F)
            for direction in (-1, +1):
                nb = np.zeros_like(loc)
                has = np.zeros(sp.shape[:-1], bool)
                if direction < 0:
                    nb[:, 1:] = loc[:, :-1]
                    has[:, 1:] = True
                else:
                    nb[:, :-1] = loc[:, 1:]
                    has[:, :-1] = True
                # NB: neighbour is transformed from its own world position; rigid transform of
                # spheres[h+-1] == loc[h+-1] (same obstacle), wp_sweep_collision_kernel.py:190-194
                half = (np.sqrt(np.sum((nb - loc) ** 2, axis=-1)) * F(0.5)).astype(F)
                inv_half = (F(1.0) / np.maximum(half, F(0.001))).astype(F)
                jump = np.zeros_like(half)
                alive = has.copy()
                for _ in range(sweep_steps):
                    alive = alive & ~(jump >= half)
                    if stats is not None:
                        stats["samples"] += int((alive & active).sum())
                    t = (F(1.0) - F(0.5) * jump * inv_half).astype(F)
                    pt = (t[..., None] * loc + (F(1.0) - t)[..., None] * nb).astype(F)
                    sdf2, n2 = sdf_fn(pt)
                    pen2 = (radj - sdf2).astype(F)
                    c2, k2 = collision_activation(pen2, eta)
                    hit2 = alive & (pen2 > 0)
                    csum = np.where(hit2, csum + c2, csum).astype(F)
                    gsum = np.where(hit2[..., None], gsum + k2[..., None] * n2, gsum).astype(F)
                    free_step = np.where(-pen2 >= F(1000.0), radj, np.maximum(-pen2, radj))
                    jump = np.where(alive, jump + np.where(pen2 > 0, pen2, free_step), jump).astype(F)
            hit = active & (csum > 0)
            gw = _quat_rotate(fq_b, gsum)
            cost[bsel] += np.where(hit, w * csum, F(0)).astype(F)
            grad[bsel, ..., :3] += np.where(hit[..., None], w * gw, F(0)).astype(F)
    if speed_dt is not None and H > 2:
        dt = max(F(speed_dt), F(1e-6))
        prev, cur, nxt = sph[:, :-2, :, :3], sph[:, 1:-1, :, :3], sph[:, 2:, :, :3]
        vel = (F(0.5) / dt * (nxt - prev)).astype(F)
        sv = np.sqrt(np.sum(vel * vel, axis=-1)).astype(F)
        d = cost[:, 1:-1]
        g = grad[:, 1:-1, :, :3]
        apply = (sv >= F(1e-3)) & (d > 0)
        with np.errstate(divide="ignore", invalid="ignore"):
            acc = ((F(1.0) / (dt * dt)) * (prev + nxt - F(2.0) * cur)).astype(F)
            nv = (vel / sv[..., None]).astype(F)
            curv = (acc / (sv * sv)[..., None]).astype(F)
            og = g - np.sum(nv * g, axis=-1, keepdims=True) * nv
            oc = curv - np.sum(nv * curv, axis=-1, keepdims=True) * nv
            new_g = (sv[..., None] * (og - d[..., None] * oc)).astype(F)
            new_d = (sv * d).astype(F)
        cost[:, 1:-1] = np.where(apply, new_d, d)
        grad[:, 1:-1, :, :3] = np.where(apply[..., None], new_g, g)
    return cost, grad


# ==========================================================================================
# Tool-pose cost
# ==========================================================================================

def _quat_mul_xyzw(a, b):
    """wp.mul(quat, quat), quats stored x,y,z,w."""
    ax, ay, az, aw = [a[..., i] for i in range(4)]
    bx, by, bz, bw = [b[..., i] for i in range(4)]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1).astype(F)


def tool_pose_cost(link_pos, link_quat, goal_pos, goal_quat, idxs_goal, weight,
                   terminal_axes_w, non_terminal_axes_w, terminal_tol, non_terminal_tol,
                   use_lie_group=False):
    """Goalset pose cost per tool frame.  cost/wp_tool_pose.py:66-204 (position / axis-angle),
    :208-295 (Lie), :457-692 (kernel), project_distance_to_goal = 0 only.

    link_pos [B,H,L,3], link_quat [B,H,L,4] wxyz; goal_* [G,L,n_goalset,3/4]; idxs_goal [B];
    weight (2,) = (w_pos, w_rot); *_axes_w [L,6]; *_tol [L,2].
    Returns cost [B,H,2L] (pos,rot interleaved), g_pos [B,H,L,3], g_quat [B,H,L,4] (wxyz quaternion-rate
    gradient = q (x) (omega,0), no 1/2 :113-126), goalset index [B,H,L], pos_err, rot_err [B,H,L]."""
    lp, lq = np.asarray(link_pos, F), np.asarray(link_quat, F)
    B, H, L, _ = lp.shape
    gp, gq = np.asarray(goal_pos, F), np.asarray(goal_quat, F)
    nG = gp.shape[2]
    wp_, wr_ = F(weight[0]), F(weight[1])
    term = np.zeros((H,), bool)
    term[:] = True
    if H > 1:
        term[:H - 1] = False
    axes = np.where(term[:, None, None], np.asarray(terminal_axes_w, F)[None], np.asarray(non_terminal_axes_w, F)[None])  # [H,L,6]
    tol = np.where(term[:, None, None], np.asarray(terminal_tol, F)[None], np.asarray(non_terminal_tol, F)[None]) ** F(2.0)
    cq = np.concatenate([lq[..., 1:], lq[..., :1]], axis=-1)      # xyzw
    best = None
    for g in range(nG):
        goal_p = gp[np.asarray(idxs_goal).astype(np.int64), :, g][:, None]            # [B,1,L,3]
        goal_q = gq[np.asarray(idxs_goal).astype(np.int64), :, g][:, None]
        goal_q = np.concatenate([goal_q[..., 1:], goal_q[..., :1]], axis=-1)
        delta = (lp - goal_p).astype(F)
        wd = delta * axes[None, ..., :3]
        pd = (F(0.5) * wp_ * np.sum(wd * wd, axis=-1)).astype(F)
        pg = (wp_ * axes[None, ..., :3] * axes[None, ..., :3] * delta).astype(F)
        z = pd < tol[None, ..., 0]
        pd = np.where(z, F(0), pd)
        pg = np.where(z[..., None], F(0), pg)
        inv_goal = goal_q * np.array([-1, -1, -1, 1], F)
        qd = _quat_mul_xyzw(cq, np.broadcast_to(inv_goal, cq.shape))
        if not use_lie_group:
            v = (axes[None, ..., 3:] * qd[..., :3]).astype(F)
            vl = np.sqrt(np.sum(v * v, axis=-1)).astype(F)
            ang = (F(2.0) * np.arctan2(vl, np.abs(qd[..., 3]))).astype(F)
            if wr_ == 0:
                ang = np.zeros_like(ang)
            with np.errstate(divide="ignore", invalid="ignore"):
                axis = np.where((vl < F(1e-15))[..., None], F(0), v / vl[..., None]).astype(F)
            om = (ang[..., None] * axis).astype(F)
            rd = (wr_ * np.sum(om * om, axis=-1)).astype(F)
            sf = np.where(qd[..., 3] < 0, F(-2.0), F(2.0))
            rg = (sf[..., None] * wr_ * om).astype(F)
            angle_out = ang
        else:
            qd = np.where((qd[..., 3] < 0)[..., None], -qd, qd)
            v = qd[..., :3]
            wq = qd[..., 3]
            vn = np.sqrt(np.sum(v * v, axis=-1)).astype(F)
            ha = np.arctan2(vn, np.abs(wq)).astype(F)
            if wr_ == 0:
                ha = np.zeros_like(ha)
            ga = (F(2.0) * ha).astype(F)
            with np.errstate(divide="ignore", invalid="ignore"):
                sinc = (ga / (F(2.0) * np.sin(ha))).astype(F)
                corr = (F(1.0) + vn * vn / (F(6.0) * wq * wq)).astype(F)
            tv = np.where((vn < F(1e-10))[..., None], F(2.0) * v,
                          np.where((np.abs(ha) < F(1e-15))[..., None], F(2.0) * v * corr[..., None], sinc[..., None] * v)).astype(F)
            wt = (axes[None, ..., 3:] * tv).astype(F)
            angle_out = np.sqrt(np.sum(wt * wt, axis=-1)).astype(F)
            rd = (wr_ * np.sum(wt * wt, axis=-1)).astype(F)
            rg = (F(2.0) * wr_ * wt).astype(F)
        z = rd < tol[None, ..., 1]
        rd = np.where(z, F(0), rd)
        rg = np.where(z[..., None], F(0), rg)
        tot = (pd + rd).astype(F)
        if best is None:
            best = dict(tot=tot, pd=pd, rd=rd, pg=pg, rg=rg, idx=np.zeros(tot.shape, np.int32), ang=angle_out)
        else:
            better = tot < best["tot"]
            for k, vv in (("tot", tot), ("pd", pd), ("rd", rd), ("ang", angle_out)):
                best[k] = np.where(better, vv, best[k])
            best["pg"] = np.where(better[..., None], pg, best["pg"])
            best["rg"] = np.where(better[..., None], rg, best["rg"])
            best["idx"] = np.where(better, np.int32(g), best["idx"])
    omq = np.concatenate([best["rg"], np.zeros_like(best["rg"][..., :1])], axis=-1)
    rate = _quat_mul_xyzw(cq, omq)                                            # xyzw
    g_quat = np.concatenate([rate[..., 3:], rate[..., :3]], axis=-1).astype(F)  # wxyz
    cost = np.stack([best["pd"], best["rd"]], axis=-1).reshape(B, H, 2 * L).astype(F)
    with np.errstate(invalid="ignore"):
        pos_err = np.where(wp_ > 0, np.sqrt(F(2.0) * best["pd"] / wp_), F(0)).astype(F)
    return cost, best["pg"].astype(F), g_quat, best["idx"], pos_err, best["ang"].astype(F)


# ==========================================================================================
# C-space costs
# ==========================================================================================

def _shrink(lo, hi, act):
    r = hi - lo
    return (lo + act * r).astype(F), (hi - act * r).astype(F)


def _bound(x, lo, hi, w):
    """aggregate_bound_cost, cost/warp_bound_util.py:24-44,74-84: 0.5*w*delta^2, grad w*delta."""
    d = np.where(x < lo, x - lo, np.where(x > hi, x - hi, F(0))).astype(F)
    return (F(0.5) * w * d * d).astype(F), (w * d).astype(F)


def cspace_position_cost(pos, limits_p, weight, activation, target=None, idxs_target=None,
                         target_weight=0.0, target_dof_weight=None):
    """IK-style c-space cost (cost_type POSITION), effort/"implied velocity" terms off.

    cost/wp_cspace_position.py:232-362 with tau weight = 0, state_dt = 0.
    pos [B,H,D]; limits_p [2,D]; weight (2,), activation (2,).  Returns cost [B,H,D], grad_p."""
    x = np.asarray(pos, F)
    D = x.shape[-1]
    lo, hi = _shrink(np.asarray(limits_p, F)[0], np.asarray(limits_p, F)[1], F(activation[0]))
    c, g = _bound(x, lo[None, None], hi[None, None], F(weight[0]))
    tw = F(target_weight) * (np.ones(D, F) if target_dof_weight is None else np.asarray(target_dof_weight, F))
    if target is not None and np.any(tw > 0):
        tgt = np.asarray(target, F)[np.asarray(idxs_target).astype(np.int64)][:, None, :]
        e = (x - tgt).astype(F)
        on = tw > 0
        c = c + np.where(on, tw * e * e, F(0))
        g = g + np.where(on, F(2.0) * tw * e, F(0))
    return c.astype(F), g.astype(F)


def cspace_state_cost(pos, vel, acc, jerk, dt, limits, weight, activation, reg_weight,
                      retime_weights=True, retime_reg=True, effort=None,
                      target=None, idxs_target=None, target_weight=0.0, non_terminal_factor=1.0,
                      target_dof_weight=None):
    """Trajectory c-space cost (cost_type STATE): bound hinge^2 on p/v/a/j(/tau), L2 smoothness on
    v/a/j(/tau), energy term, optional c-space target.  cost/wp_cspace_state.py:21-285.

    pos.. [B,H,D]; dt [B]; limits = dict(p,v,a,j,tau -> [2,D]); weight (5,), activation (5,), reg_weight (5,).
    Returns cost [B,H,D], (grad_p, grad_v, grad_a, grad_j, grad_tau)."""
    p, v, a, j = [np.asarray(t, F) for t in (pos, vel, acc, jerk)]
    B, H, D = p.shape
    tau = np.zeros_like(p) if effort is None else np.asarray(effort, F)
    dtb = np.asarray(dt, F).reshape(B, 1, 1)
    wb = [np.full((B, 1, 1), F(weight[i]), F) for i in range(5)]
    wr = [np.full((B, 1, 1), F(reg_weight[i]), F) for i in range(5)]
    if retime_weights:
        wb[1] = dtb * wb[1]
        wb[2] = np.power(dtb, F(2.0)) * wb[2]
        wb[3] = np.power(dtb, F(3.0)) * wb[3]
    if retime_reg:
        wr[0] = dtb * wr[0]
        wr[1] = np.power(dtb, F(2.0)) * wr[1]
        wr[2] = np.power(dtb, F(3.0)) * wr[2]
        wr[4] = dtb * wr[4]
    cost = np.zeros_like(p)
    grads = []
    for i, (x, key) in enumerate(((p, "p"), (v, "v"), (a, "a"), (j, "j"), (tau, "tau"))):
        lim = np.asarray(limits[key], F)
        lo, hi = _shrink(lim[0], lim[1], F(activation[i]))
        c, g = _bound(x, lo[None, None], hi[None, None], wb[i])
        cost = (cost + c).astype(F)
        grads.append(g.astype(F))
    tw = np.full((H,), F(target_weight), F)
    tw[:H - 1] = tw[:H - 1] * F(non_terminal_factor)
    if target is not None:
        dofw = np.ones(D, F) if target_dof_weight is None else np.asarray(target_dof_weight, F)
        on = (tw > 0)[None, :, None]
        twd = (tw[None, :, None] * dofw[None, None, :]).astype(F)
        tgt = np.asarray(target, F)[np.asarray(idxs_target).astype(np.int64)][:, None, :]
        e = (p - tgt).astype(F)
        cost = (cost + np.where(on, twd * e * e, F(0))).astype(F)
        grads[0] = (grads[0] + np.where(on, F(2.0) * twd * e, F(0))).astype(F)
    for i, x in ((1, v), (2, a), (3, j), (4, tau)):          # squared L2 regularisation :231-257
        wv = (wr[i - 1] * x).astype(F)
        cost = (cost + F(0.5) * wv * x).astype(F)
        grads[i] = (grads[i] + wv).astype(F)
    en_on = wr[4] > 0
    ce = (tau * v * dtb).astype(F)                            # energy :259-275, warp_bound_util.py:87-100
    cost = (cost + np.where(en_on, wr[4] * ce * ce, F(0))).astype(F)
    grads[4] = (grads[4] + np.where(en_on, F(2.0) * wr[4] * ce * v * dtb, F(0))).astype(F)
    grads[1] = (grads[1] + np.where(en_on, F(2.0) * wr[4] * ce * tau * dtb, F(0))).astype(F)
    return cost, tuple(grads)


# ==========================================================================================
# The rollout cost + gradient evaluation (what the fused kernel computes in one launch)
# ==========================================================================================

def rollout_cost_grad(rm, q, cfg, world_cuboid=None, world_voxel=None, goal_pos=None, goal_quat=None,
                      idxs_goal=None, env_query_idx=None, vel=None, acc=None, jerk=None, dt=None,
                      cspace_target=None, idxs_cspace_target=None, cspace_target_dof_weight=None):
    """One rollout cost+gradient evaluation for q [B,H,D].

    = RobotRollout.evaluate_action + sum + backward(ones) of
      optim/components/gradient_opt_core.py:445-480, i.e. (Appendix A "Backward scaling"):
        grad_q = FK_backward(g_sph_self + g_sph_scene, g_pos_pose, g_quat_pose) + g_cspace_position
    cfg: dict with keys
        self_weight, scene_weight, scene_eta, sweep(bool), speed_metric(bool),
        pose_weight(2,), pose_terminal_axes[L,6], pose_non_terminal_axes[L,6], pose_terminal_tol[L,2],
        pose_non_terminal_tol[L,2], pose_lie(bool),
        cspace_type ("position"|"state"), cspace_weight, cspace_activation, cspace_reg (state only),
        retime_weights, retime_reg, cspace_target_weight, cspace_non_terminal_weight_factor
    cspace_target [n,D] / idxs_cspace_target [B] / cspace_target_dof_weight [D]: the c-space target term
    (wp_cspace_state.py:84-89,220-226; wp_cspace_position.py).  Several link-sphere configurations
    (rm.link_spheres [n_cfg,S,4]) are selected per seed through env_query_idx like the world.
    Returns dict(cost[B] (sum over h and terms), cost_bh[B,H], grad_q[B,H,D], + per-term outputs)."""
    q = np.asarray(q, F)
    B, H, D = q.shape
    N = B * H
    multi_cfg = np.asarray(rm.link_spheres).ndim == 3 and np.asarray(rm.link_spheres).shape[0] > 1
    cfg_idx = (np.zeros(B, np.int64) if env_query_idx is None else np.asarray(env_query_idx)) if multi_cfg else None
    cum, sph, lpos, lquat = fk_forward(rm, q.reshape(N, D), cfg_idx, H)
    out = {"cumul": cum, "spheres": sph.reshape(B, H, -1, 4), "link_pos": lpos.reshape(B, H, -1, 3),
           "link_quat": lquat.reshape(B, H, -1, 4)}
    g_sph = np.zeros_like(sph)
    cost_bh = np.zeros((B, H), F)
    if cfg.get("self_weight", 0) > 0:
        c, g, k = self_collision(sph, rm.sphere_padding, rm.collision_pairs, cfg["self_weight"])
        out["self_cost"], out["self_grad"], out["self_pair"] = c.reshape(B, H), g.reshape(B, H, -1, 4), k.reshape(B, H)
        g_sph += g
        cost_bh += c.reshape(B, H)
    if cfg.get("scene_weight", 0) > 0 and (world_cuboid is not None or world_voxel is not None):
        sdt = None
        if cfg.get("sweep") and cfg.get("speed_metric"):
            sdt = float(np.asarray(dt).reshape(-1)[0])      # one dt for the batch, wp_speed_metric.py:54
        c, g = scene_collision(out["spheres"], cfg["scene_weight"], cfg.get("scene_eta", 0.0), world_cuboid,
                               world_voxel, env_query_idx, sweep=bool(cfg.get("sweep")), speed_dt=sdt)
        out["scene_cost"], out["scene_grad"] = c, g
        g_sph += g.reshape(N, -1, 4)
        cost_bh += np.sum(c, axis=-1)
    g_pos = g_quat = None
    if goal_pos is not None and cfg.get("pose_weight") is not None:
        L = lpos.shape[1]
        ones6, zeros2 = np.ones((L, 6), F), np.zeros((L, 2), F)
        c, g_pos, g_quat, gi, pe, re = tool_pose_cost(
            out["link_pos"], out["link_quat"], goal_pos, goal_quat,
            np.zeros(B, np.int64) if idxs_goal is None else idxs_goal, cfg["pose_weight"],
            cfg.get("pose_terminal_axes", ones6), cfg.get("pose_non_terminal_axes", ones6),
            cfg.get("pose_terminal_tol", zeros2), cfg.get("pose_non_terminal_tol", zeros2),
            use_lie_group=bool(cfg.get("pose_lie", False)))
        out.update(pose_cost=c, pose_grad_pos=g_pos, pose_grad_quat=g_quat, pose_goalset_idx=gi,
                   pose_pos_err=pe, pose_rot_err=re)
        cost_bh += np.sum(c, axis=-1)
        g_pos, g_quat = g_pos.reshape(N, L, 3), g_quat.reshape(N, L, 4)
    gq = fk_backward(rm, cum, g_sph, g_pos, g_quat, cfg_idx, H).reshape(B, H, D)
    ctype = cfg.get("cspace_type")
    tgt_w = float(cfg.get("cspace_target_weight", 0.0)) if cspace_target is not None else 0.0
    tgt_idx = np.zeros(B, np.int64) if idxs_cspace_target is None else idxs_cspace_target
    if ctype == "position":
        c, gp = cspace_position_cost(q, rm.position_limits, cfg["cspace_weight"], cfg["cspace_activation"],
                                     target=cspace_target if tgt_w > 0 else None, idxs_target=tgt_idx,
                                     target_weight=tgt_w, target_dof_weight=cspace_target_dof_weight)
        out["cspace_cost"], out["cspace_grad_p"] = c, gp
        cost_bh += np.sum(c, axis=-1)
        gq = (gq + gp).astype(F)
    elif ctype == "state":
        z = np.zeros_like(q)
        lim = dict(p=rm.position_limits, v=rm.velocity_limits, a=rm.acceleration_limits, j=rm.jerk_limits,
                   tau=rm.effort_limits)
        c, gs = cspace_state_cost(q, z if vel is None else vel, z if acc is None else acc,
                                  z if jerk is None else jerk, np.ones(B, F) if dt is None else dt, lim,
                                  cfg["cspace_weight"], cfg["cspace_activation"], cfg["cspace_reg"],
                                  cfg.get("retime_weights", True), cfg.get("retime_reg", True),
                                  target=cspace_target if tgt_w > 0 else None, idxs_target=tgt_idx, target_weight=tgt_w,
                                  non_terminal_factor=float(cfg.get("cspace_non_terminal_weight_factor", 1.0)),
                                  target_dof_weight=cspace_target_dof_weight)
        out["cspace_cost"], out["cspace_grads"] = c, gs
        cost_bh += np.sum(c, axis=-1)
        gq = (gq + gs[0]).astype(F)
    out["grad_q"] = gq
    out["cost_bh"] = cost_bh
    out["cost"] = np.sum(cost_bh, axis=1).astype(F)
    return out
