"""CPU ORACLE for RNEA inverse dynamics and its adjoint  --  TEST INFRASTRUCTURE, NOT PRODUCT.

SURVEY.md section 8(f) rank 3: tau = RNEA(q, qd, qdd) (the torque that feeds the effort terms of the STATE c-space cost)
and d loss / d (q, qd, qdd) from d loss / d tau.  Only tests/ and __graft_entry__.smoke() may import this module.

Float32 numpy restatement of the reference's CUDA kernels in their serial (threads_per_batch = 1) order, vectorised
over the batch (paths relative to curobo/_src/curobolib/kernels/dynamics/):
  rnea_forward_kernel   rnea_forward_kernel.cuh:54-285
  rnea_backward_kernel  rnea_backward_kernel.cuh:60-460
  spatial algebra       spatial_algebra.cuh (Featherstone order: angular first, [w; v]); rnea_helpers.cuh:25-95
Pinned on the GPU box against the reference's own kernels compiled into oracle/_ref, and here by finite differences
of the forward pass and by the physics identities tau(q, 0, 0) = gravity torque, tau linear in qdd.
"""
from __future__ import annotations

import numpy as np

F = np.float32
FIXED, X_PRISM, Y_PRISM, Z_PRISM, X_ROT, Y_ROT, Z_ROT = -1, 0, 1, 2, 3, 4, 5


def s_index(jt: int) -> int:
    """rnea_helpers.cuh:13-18: revolute -> angular slot 0..2, prismatic -> linear slot 3..5."""
    return jt - X_ROT if jt >= X_ROT else 3 + jt


def tree_levels(link_map):
    """CSR (level_starts, level_links) of the kinematic tree by depth (kp.link_level_offsets / link_level_data)."""
    nl = len(link_map)
    depth = np.zeros(nl, np.int32)
    for k in range(1, nl):
        depth[k] = depth[int(link_map[k])] + 1
    order = np.argsort(depth, kind="stable").astype(np.int16)
    starts = np.zeros(depth.max() + 2, np.int16)
    for d in depth:
        starts[d + 1] += 1
    return np.cumsum(starts).astype(np.int16), order


def local_Rp(ft, jt, angle):
    """compute_local_Rp (rnea_helpers.cuh:25-95): ft [12] row-major 3x4, angle [B] -> R [B,9] (row-major of
    fixed.R * J(angle)), p [B,3] (fixed translation + prismatic offset)."""
    B = angle.shape[0]
    f = np.asarray(ft, F).reshape(3, 4)
    R = np.broadcast_to(f[:, :3].reshape(9), (B, 9)).astype(F).copy()
    p = np.broadcast_to(f[:, 3], (B, 3)).astype(F).copy()
    if jt == FIXED:
        return R, p
    if jt >= X_ROT:
        s, c = np.sin(angle).astype(F), np.cos(angle).astype(F)
        ax = jt - X_ROT
        for r in range(3):
            x, y, z = f[r, 0], f[r, 1], f[r, 2]
            if ax == 0:
                R[:, 3 * r + 1] = (c * y + s * z).astype(F)
                R[:, 3 * r + 2] = (-s * y + c * z).astype(F)
            elif ax == 1:
                R[:, 3 * r + 0] = (c * x - s * z).astype(F)
                R[:, 3 * r + 2] = (s * x + c * z).astype(F)
            else:
                R[:, 3 * r + 0] = (c * x + s * y).astype(F)
                R[:, 3 * r + 1] = (-s * x + c * y).astype(F)
    else:
        p = (p + f[:, jt][None, :] * angle[:, None]).astype(F)
    return R, p


def _cross(a, b):
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], 1).astype(F)


def _Rt(R, x):  # R^T x with R row-major [B,9]   (spatial_Xv uses R[0]*w0 + R[3]*w1 + R[6]*w2 ...)
    return np.stack([R[:, 0] * x[:, 0] + R[:, 3] * x[:, 1] + R[:, 6] * x[:, 2], R[:, 1] * x[:, 0] + R[:, 4] * x[:, 1] + R[:, 7] * x[:, 2],
                     R[:, 2] * x[:, 0] + R[:, 5] * x[:, 1] + R[:, 8] * x[:, 2]], 1).astype(F)


def _R(R, x):  # R x
    return np.stack([R[:, 0] * x[:, 0] + R[:, 1] * x[:, 1] + R[:, 2] * x[:, 2], R[:, 3] * x[:, 0] + R[:, 4] * x[:, 1] + R[:, 5] * x[:, 2],
                     R[:, 6] * x[:, 0] + R[:, 7] * x[:, 1] + R[:, 8] * x[:, 2]], 1).astype(F)


def spatial_Xv(R, p, v):
    """Motion transform parent -> child (spatial_algebra.cuh:22-40): w' = R^T w, v' = R^T (v + w x p)."""
    w, vl = v[:, :3], v[:, 3:]
    u = (vl + _cross(w, p)).astype(F)
    return np.concatenate([_Rt(R, w), _Rt(R, u)], 1).astype(F)


def spatial_XTf(R, p, f):
    """Force transform child -> parent (spatial_algebra.cuh:42-66): f' = R f, n' = R n + p x (R f)."""
    n, fl = f[:, :3], f[:, 3:]
    Rfl, Rn = _R(R, fl), _R(R, n)
    return np.concatenate([(Rn + _cross(p, Rfl)).astype(F), Rfl], 1).astype(F)


def motion_cross(a, b):
    """crm(a) b (spatial_algebra.cuh:185-198)."""
    aw, av, bw, bv = a[:, :3], a[:, 3:], b[:, :3], b[:, 3:]
    return np.concatenate([_cross(aw, bw), (_cross(av, bw) + _cross(aw, bv)).astype(F)], 1).astype(F)


def force_cross(v, f):
    """crf(v) f (spatial_algebra.cuh:112-127)."""
    w, vl, n, fl = v[:, :3], v[:, 3:], f[:, :3], f[:, 3:]
    return np.concatenate([(_cross(w, n) + _cross(vl, fl)).astype(F), _cross(w, fl)], 1).astype(F)


# Joint-axis specialised operators, restated entry by entry (spatial_algebra.cuh:68-110, 200-226).  NB for prismatic
# joints the reference's motion_cross_S writes crm(v) S into the ANGULAR slots 0..2 (the generic product puts it in
# the linear slots); parity means reproducing that.
_MCS = {0: ((1, 2, 1), (2, 1, -1), (4, 5, 1), (5, 4, -1)), 1: ((0, 2, -1), (2, 0, 1), (3, 5, -1), (5, 3, 1)),
        2: ((0, 1, 1), (1, 0, -1), (3, 4, 1), (4, 3, -1)), 3: ((1, 2, 1), (2, 1, -1)), 4: ((0, 2, -1), (2, 0, 1)),
        5: ((0, 1, 1), (1, 0, -1))}
_CRF = {0: ((1, 2, -1), (2, 1, 1), (4, 5, -1), (5, 4, 1)), 1: ((0, 2, 1), (2, 0, -1), (3, 5, 1), (5, 3, -1)),
        2: ((0, 1, -1), (1, 0, 1), (3, 4, -1), (4, 3, 1)), 3: ((1, 5, -1), (2, 4, 1)), 4: ((0, 5, 1), (2, 3, -1)),
        5: ((0, 4, -1), (1, 3, 1))}
_CRM = {0: _CRF[0], 1: _CRF[1], 2: _CRF[2], 3: ((4, 2, -1), (5, 1, 1)), 4: ((3, 2, 1), (5, 0, -1)), 5: ((3, 1, -1), (4, 0, 1))}


def motion_cross_S(v, s, alpha):
    out = np.zeros_like(v)
    for o, i, sg in _MCS[s]:
        out[:, o] = (F(sg) * v[:, i] * alpha).astype(F)
    return out


def _dot_S(table, a, b, s):
    r = np.zeros(a.shape[0], F)
    for i, j, sg in table[s]:
        r = (r + F(sg) * a[:, i] * b[:, j]).astype(F)
    return r


def dot_crf_S(a, b, s):
    return _dot_S(_CRF, a, b, s)


def dot_crm_S(a, b, s):
    return _dot_S(_CRM, a, b, s)


def force_cross_S_add(res, s, alpha, b):
    """res += crf(S alpha) b (spatial_algebra.cuh:228-252)."""
    e = np.zeros_like(b)
    e[:, s] = alpha
    return (res + force_cross(e, b)).astype(F)


def inertia_times(mc, inertia, u):
    """Spatial inertia about the link origin times a motion vector (spatial_algebra.cuh:129-163).
    mc = (cx, cy, cz, m); inertia = (ixx, iyy, izz, ixy, ixz, iyz, pad, pad) at the CoM."""
    c, m = np.asarray(mc[:3], F), F(mc[3])
    ixx, iyy, izz, ixy, ixz, iyz = (F(x) for x in inertia[:6])
    w, vl = u[:, :3], u[:, 3:]
    cb = np.broadcast_to(c, w.shape)
    h = (vl + _cross(w, cb)).astype(F)
    Iw = np.stack([ixx * w[:, 0] + ixy * w[:, 1] + ixz * w[:, 2], ixy * w[:, 0] + iyy * w[:, 1] + iyz * w[:, 2],
                   ixz * w[:, 0] + iyz * w[:, 1] + izz * w[:, 2]], 1).astype(F)
    return np.concatenate([(Iw + m * _cross(cb, h)).astype(F), (m * h).astype(F)], 1).astype(F)


def rnea_forward(q, qd, qdd, fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map, link_map,
                 joint_offset_map, gravity, f_ext=None):
    """tau [B,D] and the forward cache (v, a, f per link, [B,nl,6] each).  gravity [6] spatial (0,0,0, gx,gy,gz) as the
    reference passes it (the base's fictitious acceleration, i.e. -g for a z-up world)."""
    q, qd, qdd = (np.asarray(x, F) for x in (q, qd, qdd))
    B, D = q.shape
    nl = len(link_map)
    starts, order = tree_levels(link_map)
    joff = np.asarray(joint_offset_map, F).reshape(nl, 2)
    v = np.zeros((B, nl, 6), F)
    a = np.zeros((B, nl, 6), F)
    g = np.broadcast_to(np.asarray(gravity, F), (B, 6))
    Rp = {}
    for k in order:                                   # rnea_forward_kernel.cuh:104-168 (level order == depth order)
        k = int(k)
        jt, ji, par = int(joint_map_type[k]), int(joint_map[k]), int(link_map[k])
        root = par < 0 or par == k
        qe = qde = qdde = np.zeros(B, F)
        if jt != FIXED and ji >= 0:
            mul = F(joff[k, 0])
            qe = (mul * q[:, ji] + F(joff[k, 1])).astype(F)
            qde = (mul * qd[:, ji]).astype(F)
            qdde = (mul * qdd[:, ji]).astype(F)
        R, p = local_Rp(fixed_transforms[k], jt, qe)
        Rp[k] = (R, p)
        if root:
            vk = np.zeros((B, 6), F)
            ak = spatial_Xv(R, p, g)
        else:
            vk = spatial_Xv(R, p, v[:, par])
            ak = spatial_Xv(R, p, a[:, par])
        if jt != FIXED:
            s = s_index(jt)
            vk[:, s] = (vk[:, s] + qde).astype(F)
            ak[:, s] = (ak[:, s] + qdde).astype(F)
            ak = (ak + motion_cross_S(vk, s, qde)).astype(F)
        v[:, k], a[:, k] = vk, ak
    f = np.zeros((B, nl, 6), F)
    for k in range(nl):                               # :171-199
        Ia = inertia_times(link_masses_com[k], link_inertias[k], a[:, k])
        Iv = inertia_times(link_masses_com[k], link_inertias[k], v[:, k])
        f[:, k] = (Ia + force_cross(v[:, k], Iv)).astype(F)
        if f_ext is not None:
            f[:, k] = (f[:, k] - np.asarray(f_ext, F)[:, k]).astype(F)
    tau = np.zeros((B, D), F)
    # :203-253: reverse LEVEL order with ascending index inside a level, like the serial kernel
    for lv in range(len(starts) - 2, -1, -1):
        for k in order[starts[lv]:starts[lv + 1]]:
            k = int(k)
            jt, ji, par = int(joint_map_type[k]), int(joint_map[k]), int(link_map[k])
            root = par < 0 or par == k
            if jt != FIXED and ji >= 0:
                tau[:, ji] = (tau[:, ji] + F(joff[k, 0]) * f[:, k, s_index(jt)]).astype(F)
            if not root:
                R, p = Rp[k]
                f[:, par] = (f[:, par] + spatial_XTf(R, p, f[:, k])).astype(F)
    return tau, dict(v=v, a=a, f=f, Rp=Rp, levels=(starts, order))


def rnea_backward(grad_tau, q, qd, cache, fixed_transforms, link_masses_com, link_inertias, joint_map_type, joint_map,
                  link_map, joint_offset_map, gravity):
    """(grad_q, grad_qd, grad_qdd) [B,D] from grad_tau, using the forward cache (rnea_backward_kernel.cuh:100-455)."""
    q, qd, gt = np.asarray(q, F), np.asarray(qd, F), np.asarray(grad_tau, F)
    B, D = q.shape
    nl = len(link_map)
    starts, order = cache["levels"]
    joff = np.asarray(joint_offset_map, F).reshape(nl, 2)
    v, a, Rp = cache["v"], cache["a"], cache["Rp"]
    sf = cache["f"].copy()
    g = np.broadcast_to(np.asarray(gravity, F), (B, 6))
    gq, gqd, gqdd = np.zeros((B, D), F), np.zeros((B, D), F), np.zeros((B, D), F)

    crf_S, crm_S = dot_crf_S, dot_crm_S

    # pass 1 (root -> leaves): f_bar[k] = S grad_tau + X f_bar[parent]; grad_q += mult * <X f_bar[parent], f[k]>_crf
    for lv in range(len(starts) - 1):
        for k in order[starts[lv]:starts[lv + 1]]:
            k = int(k)
            jt, ji, par = int(joint_map_type[k]), int(joint_map[k]), int(link_map[k])
            root = par < 0 or par == k
            fk = sf[:, k].copy()
            fbar = np.zeros((B, 6), F)
            mul = F(1)
            if jt != FIXED and ji >= 0:
                mul = F(joff[k, 0])
                fbar[:, s_index(jt)] = (fbar[:, s_index(jt)] + mul * gt[:, ji]).astype(F)
            if not root:
                R, p = Rp[k]
                Xf = spatial_Xv(R, p, sf[:, par])
                fbar = (fbar + Xf).astype(F)
                if jt != FIXED and ji >= 0:
                    gq[:, ji] = (gq[:, ji] + mul * crf_S(Xf, fk, s_index(jt))).astype(F)
            sf[:, k] = fbar
    abar = np.zeros((B, nl, 6), F)
    vbar = np.zeros((B, nl, 6), F)
    # pass 2 (leaves -> root)
    for lv in range(len(starts) - 2, -1, -1):
        for k in order[starts[lv]:starts[lv + 1]]:
            k = int(k)
            jt, ji, par = int(joint_map_type[k]), int(joint_map[k]), int(link_map[k])
            root = par < 0 or par == k
            mc, inn = link_masses_com[k], link_inertias[k]
            vk, fbar = v[:, k], sf[:, k]
            ab = (abar[:, k] + inertia_times(mc, inn, fbar)).astype(F)
            Iv = inertia_times(mc, inn, vk)
            vb = (vbar[:, k] - force_cross(fbar, Iv)).astype(F)
            vb = (vb - inertia_times(mc, inn, motion_cross(vk, fbar))).astype(F)
            mul = F(1)
            R, p = Rp[k]
            if jt != FIXED and ji >= 0:
                mul = F(joff[k, 0])
                s = s_index(jt)
                qdk = (mul * qd[:, ji]).astype(F)
                gqdd[:, ji] = (gqdd[:, ji] + mul * ab[:, s]).astype(F)
                gqd[:, ji] = (gqd[:, ji] - mul * force_cross(vk, ab)[:, s]).astype(F)
                vb = force_cross_S_add(vb, s, qdk, ab)
            if not root:
                abar[:, par] = (abar[:, par] + spatial_XTf(R, p, ab)).astype(F)
                if jt != FIXED and ji >= 0:
                    gq[:, ji] = (gq[:, ji] - mul * crm_S(ab, spatial_Xv(R, p, a[:, par]), s_index(jt))).astype(F)
            elif jt != FIXED and ji >= 0:
                gq[:, ji] = (gq[:, ji] - mul * crm_S(ab, spatial_Xv(R, p, g), s_index(jt))).astype(F)
            if jt != FIXED and ji >= 0:
                gqd[:, ji] = (gqd[:, ji] + mul * vb[:, s_index(jt)]).astype(F)
            if not root:
                vbar[:, par] = (vbar[:, par] + spatial_XTf(R, p, vb)).astype(F)
                if jt != FIXED and ji >= 0:
                    gq[:, ji] = (gq[:, ji] - mul * crm_S(vb, spatial_Xv(R, p, v[:, par]), s_index(jt))).astype(F)
    return gq, gqd, gqdd
