"""CPU ORACLE for the B-spline knot -> state kernel and its adjoint  --  TEST INFRASTRUCTURE, NOT PRODUCT.

SURVEY.md section 8(f) rank 1: the step immediately before FK / after FK-backward on the trajopt / MPC path.
Only tests/ and __graft_entry__.smoke() may import this module; curobo_b200/ never does.

Float32 numpy restatement of the reference's CUDA arithmetic (MATRIX basis backend, which is what both the
pybind launcher and the cuda.core backend instantiate: backends/pybind/trajectory_kernel_launch.cu:259,
backends/cuda_core_backend/trajectory.py:71,163,252).  Paths below are relative to
curobo/_src/curobolib/kernels/trajectory/bspline/.

Pinning status: the reference ships no golden vectors for this kernel (its tests are no-NaN + gradcheck,
tests/_src/transition/test_transition_gradients.py:390-600); the oracle is pinned on the GPU box against the
reference's own kernels compiled from /root/reference into oracle/_ref (tests/test_gpu_bspline.py) and by the
algebraic properties the reference relies on (partition of unity, boundary state reproduction, adjointness).
"""
from __future__ import annotations

import numpy as np

F = np.float32
FP32_PRECISION = F(1e-6)  # common/curobo_constants.h:28

# basis/bspline_basis_matrix.cuh:14-39 -- rows = basis function, columns = powers t^n .. t^0
_M3 = np.array([[-1, 3, -3, 1], [3, -6, 0, 4], [-3, 3, 3, 1], [1, 0, 0, 0]], np.float64) / 6.0
_M4 = np.array([[1, -4, 6, -4, 1], [-4, 12, -6, -12, 11], [6, -12, -6, 12, 11], [-4, 4, 6, 4, 1], [1, 0, 0, 0, 0]],
               np.float64) / 24.0
_M5 = np.array([[-1, 5, -10, 10, -5, 1], [5, -20, 20, 20, -50, 26], [-10, 30, 0, -60, 0, 66],
                [10, -20, -20, 20, 50, 26], [-5, 5, 10, 10, 5, 1], [1, 0, 0, 0, 0, 0]], np.float64) / 120.0
BASIS_MATRIX = {3: _M3.astype(F), 4: _M4.astype(F), 5: _M5.astype(F)}

# bspline_boundary_constraint.cuh:33-73 -- fixed (virtual) knots from a boundary state:
# fixed[i] = cp[i]*pos + cv[i]*vel*kdt + ca[i]*acc*kdt^2 + cj[i]*jerk*kdt^3   (:104-118)
FIXED_KNOT_COEFFS = {
    3: np.array([[1, 1, 1, 1], [-1, 0, 1, 2], [1 / 3, -1 / 6, 1 / 3, 11 / 6], [0, 0, 0, 0]], np.float64).astype(F),
    4: np.array([[1, 1, 1, 1, 1], [-1.5, -0.5, 0.5, 1.5, 2.5], [11 / 12, -1 / 12, -1 / 12, 11 / 12, 35 / 12],
                 [-3 / 12, 1 / 12, -1 / 12, 3 / 12, 25 / 12]], np.float64).astype(F),
    5: np.array([[1, 1, 1, 1, 1, 1], [-2, -1, 0, 1, 2, 3], [1.75, 0.25, -0.25, 0.25, 1.75, 4.25],
                 [-0.833333, 0.083333, 0.0, -0.083333, 0.833333, 3.75]], np.float64).astype(F),
}


def total_knots(n_knots: int, degree: int) -> int:
    """bspline_common.cuh:21-24"""
    return n_knots + degree + 1


def padded_horizon_for(n_knots: int, degree: int, interpolation_steps: int) -> int:
    """types/control_space.py:44-49"""
    return total_knots(n_knots, degree) * interpolation_steps + 1


def basis_rows(degree: int, t):
    """MATRIX backend: basis_k[i] = sum_j M[i][j] * tp_k[j] over the first (S-k) columns, accumulated left to
    right from 0 (common/math.cuh:73-104), with the derivative power vectors of
    basis/bspline_basis_matrix.cuh:47-151.  t: float32 array [...] -> four arrays [..., S]."""
    t = np.asarray(t, F)
    M = BASIS_MATRIX[degree]
    S = degree + 1
    one = np.ones_like(t)

    def pw(n):  # t*t*...*t, left to right
        r = t
        for _ in range(n - 1):
            r = (r * t).astype(F)
        return r

    if degree == 3:
        tp = [[pw(3), pw(2), t, one], [F(3) * t * t, F(2) * t, one], [F(6) * t, F(2) * one], [F(6) * one]]
    elif degree == 4:
        tp = [[pw(4), pw(3), pw(2), t, one], [F(4) * t * t * t, F(3) * t * t, F(2) * t, one],
              [F(12) * t * t, F(6) * t, F(2) * one], [F(24) * t, F(6) * one]]
    elif degree == 5:
        tp = [[pw(5), pw(4), pw(3), pw(2), t, one], [F(5) * t * t * t * t, F(4) * t * t * t, F(3) * t * t, F(2) * t, one],
              [F(20) * t * t * t, F(12) * t * t, F(6) * t, F(2) * one], [F(60) * t * t, F(24) * t, F(6) * one]]
    else:
        raise ValueError("bspline degree must be 3, 4 or 5")
    out = []
    for k in range(4):
        cols = [np.asarray(c, F) for c in tp[k]]
        b = np.zeros(t.shape + (S,), F)
        for i in range(S):
            acc = np.zeros_like(t)
            for j, c in enumerate(cols):
                acc = (acc + M[i, j] * c).astype(F)
            b[..., i] = acc
        out.append(b)
    return out


def _fixed_knots(degree, pos, vel, acc, jerk, kdt):
    """bspline_boundary_constraint.cuh:104-118; pos.. [..] -> [.., S]"""
    C = FIXED_KNOT_COEFFS[degree]
    kdt2 = (kdt * kdt).astype(F)
    kdt3 = (kdt * kdt * kdt).astype(F)  # bspline_context.cuh:56
    out = np.zeros(pos.shape + (degree + 1,), F)
    for i in range(degree + 1):
        out[..., i] = (((C[0, i] * pos + C[1, i] * vel * kdt).astype(F) + C[2, i] * acc * kdt2).astype(F)
                       + C[3, i] * jerk * kdt3).astype(F)
    return out


def _dot(knots, basis):
    """common/math.cuh:20-28: result = 0; result += knots[i]*basis[i] in index order."""
    r = np.zeros(knots.shape[:-1], F)
    for i in range(knots.shape[-1]):
        r = (r + knots[..., i] * basis[..., i]).astype(F)
    return r


def local_support(knots, start, goal, start_idx, goal_idx, use_implicit_goal_state, knot_dt, knot_idx, degree):
    """The S local control points of spline segment `knot_idx` for every (b, d): user knots where they exist,
    fixed knots from the start state for the first S segments, and for the tail either fixed knots from the goal
    state (implicit) or the last knot replicated (bspline_interpolation.cuh:170-228 +
    bspline_boundary_constraint.cuh:121-265).  knots [B,nk,D]; start/goal: 4-tuples of [N,D]; knot_dt [B]."""
    knots = np.asarray(knots, F)
    B, nk, D = knots.shape
    S = degree + 1
    loc = np.zeros((B, D, S), F)
    s0 = knot_idx - S
    for i in range(S):
        src = s0 + i
        if 0 <= src < nk:
            loc[:, :, i] = knots[:, src, :]
    implicit = np.asarray(use_implicit_goal_state)[goal_idx].astype(bool)  # [B]
    kdt = np.broadcast_to(np.asarray(knot_dt, F)[:, None], (B, D))
    if knot_idx < S:  # start boundary (bspline_common.cuh:27-29)
        fx = _fixed_knots(degree, *(np.asarray(x, F)[start_idx] for x in start), kdt)
        for i in range(S - knot_idx):
            loc[:, :, i] = fx[:, :, knot_idx + i]
        return loc
    fxg = None
    for b in range(B):
        if implicit[b] and knot_idx > nk - 1:  # :36-38
            if fxg is None:
                fxg = _fixed_knots(degree, *(np.asarray(x, F)[goal_idx] for x in goal), kdt)
            n = knot_idx - nk + 1
            for i in range(n):
                loc[b, :, S - n + i] = fxg[b, :, i]
        elif (not implicit[b]) and knot_idx > nk:  # :31-33
            n = knot_idx - nk
            src = loc[b, :, S - n - 1].copy()
            for i in range(n):
                loc[b, :, S - 1 - i] = src
    return loc


def bspline_forward(knots, start, goal, start_idx, goal_idx, traj_dt, use_implicit_goal_state, padded_horizon, degree,
                    interpolation_horizon=None, interpolation_dt=None):
    """interpolate_bspline_kernel (bspline_kernel.cuh:87-149) / interpolate_bspline_single_dt_kernel (:216-270)
    through interpolate_bspline_trajectory (bspline_interpolation.cuh:21-296).

    knots [B,nk,D]; start = (pos, vel, acc, jerk) each [Ns,D]; goal likewise [Ng,D]; start_idx, goal_idx int [B];
    traj_dt [Ng] and use_implicit_goal_state [Ng] are indexed through goal_idx.
    Returns pos, vel, acc, jerk [B,padded_horizon,D] and out_dt [B].

    single-dt variant: pass interpolation_horizon int [B] and interpolation_dt (scalar): every batch uses the same
    dt, its own horizon min(interpolation_horizon[b], T-1), rows beyond it keep clamped-at-the-end values."""
    knots = np.asarray(knots, F)
    B, nk, D = knots.shape
    S = degree + 1
    T = int(padded_horizon)
    start_idx = np.asarray(start_idx, np.int64)
    goal_idx = np.asarray(goal_idx, np.int64)
    pnk = total_knots(nk, degree)
    out = [np.zeros((B, T, D), F) for _ in range(4)]
    if interpolation_horizon is None:
        groups = {T - 1: np.arange(B)}
        dt_b = np.asarray(traj_dt, F)[goal_idx]
    else:
        hz = np.minimum(np.asarray(interpolation_horizon, np.int64), T - 1)
        groups = {int(h): np.nonzero(hz == h)[0] for h in np.unique(hz)}
        dt_b = np.full((B,), F(interpolation_dt), F)
    for horizon, rows in groups.items():
        steps = horizon // pnk  # :63
        kdt_rows = (np.maximum(dt_b[rows], FP32_PRECISION) * F(steps)).astype(F)  # :64
        cache = {}
        for h in range(T):
            knot_idx = (h // steps) if steps > 0 else 0
            t_mod = F(0.0)
            if steps > 0:
                t_mod = F(F(h) / F(steps)) - F(int(h / steps))  # :236-238
            if knot_idx >= pnk:
                knot_idx = pnk - 1
                t_mod = F(1.0)  # :76-80, :240
            if knot_idx not in cache:
                cache[knot_idx] = local_support(knots[rows], start, goal, start_idx[rows], goal_idx[rows],
                                                use_implicit_goal_state, kdt_rows, knot_idx, degree)
            loc = cache[knot_idx]  # [b,D,S]
            bp, bv, ba, bj = (x[None, None, :] for x in basis_rows(degree, np.asarray(t_mod, F)))
            kd = kdt_rows[:, None]
            out[0][rows, h] = _dot(loc, np.broadcast_to(bp, loc.shape))
            out[1][rows, h] = (_dot(loc, np.broadcast_to(bv, loc.shape)) / kd).astype(F)
            out[2][rows, h] = (_dot(loc, np.broadcast_to(ba, loc.shape)) / (kd * kd).astype(F)).astype(F)
            out[3][rows, h] = (_dot(loc, np.broadcast_to(bj, loc.shape)) / (kd * kd * kd).astype(F)).astype(F)
    return out[0], out[1], out[2], out[3], dt_b.copy()


def bspline_backward(grad_pos, grad_vel, grad_acc, grad_jerk, traj_dt, dt_idx, use_implicit_goal_state, n_knots, degree):
    """bspline_backward_kernel (bspline_kernel.cuh:326-373): gradient of the loss wrt the knots.

    grad_* [B,padded_horizon,D].  For knot k the contributing rows are h = (k+1+i)*steps + j, i in [0,S), j in
    [0,steps) (bspline_gradient_util.cuh:85-104), paired with basis[S-1-i] at t = j/steps
    (bspline_context.cuh:152-186).  Quirks kept as in the reference:
      * implicit goal: knots n_knots-1 (and beyond) get zero gradient (:83), since that knot is overwritten by the goal;
      * replicate: the last knot also collects the rows where it is replicated (:106-125) and the padded last row's
        POSITION gradient only (:127-147);
      * knot_dt is not clamped here (bspline_common.cuh:172-173), unlike the forward pass.
    The sum over j is a plain left-to-right sum (the reference's shuffle tree, bspline_gradient_util.cuh:34-55, is
    only correct for power-of-two interpolation_steps and differs from this by float re-association)."""
    gp, gv, ga, gj = (np.asarray(x, F) for x in (grad_pos, grad_vel, grad_acc, grad_jerk))
    B, T, D = gp.shape
    S = degree + 1
    horizon = T - 1
    pnk = total_knots(n_knots, degree)
    steps = horizon // pnk
    if steps <= 0:
        raise ValueError("interpolation_steps is 0")
    dt_idx = np.asarray(dt_idx, np.int64)
    use_goal = np.asarray(use_implicit_goal_state)[dt_idx].astype(bool)
    kdt = (np.asarray(traj_dt, F)[dt_idx] * F(steps)).astype(F)[:, None]
    kdt2 = (kdt * kdt).astype(F)
    kdt3 = (kdt * kdt * kdt).astype(F)
    ext = pnk * steps
    out = np.zeros((B, n_knots, D), F)
    for k in range(n_knots):
        zero_rows = use_goal & (k >= n_knots - 1)
        repl_rows = (~use_goal) & (k == n_knots - 1)
        tot = np.zeros((B, D), F)
        for j in range(steps):
            g = np.zeros((4, B, D, S), F)
            for i in range(S):
                oh = (k + 1) * steps + j + i * steps
                if oh < ext:
                    for a, arr in enumerate((gp, gv, ga, gj)):
                        g[a, :, :, i] = arr[:, oh, :]
            g[:, zero_rows] = 0
            if repl_rows.any():
                gr = g[:, repl_rows].copy()
                for i in range(1, S):
                    for x in range(i):
                        gr[..., x] = (gr[..., x] + gr[..., i]).astype(F)
                if j == 0:
                    gr[0] = (gr[0] + gp[repl_rows, horizon, :][..., None]).astype(F)
                g[:, repl_rows] = gr
            h_idx = (k + degree) * steps + j
            t_mod = F(F(h_idx) / F(steps)) - F(int(h_idx / steps))  # bspline_common.cuh:175
            basis = basis_rows(degree, np.asarray(t_mod, F))
            sums = []
            for a in range(4):
                r = np.zeros((B, D), F)
                for i in range(S):  # dot_product_reverse, common/math.cuh:31-38
                    r = (r + g[a, :, :, i] * basis[a][S - 1 - i]).astype(F)
                sums.append(r)
            val = (((sums[0] + (sums[1] / kdt).astype(F)).astype(F) + (sums[2] / kdt2).astype(F)).astype(F)
                   + (sums[3] / kdt3).astype(F)).astype(F)
            tot = (tot + val).astype(F)
        out[:, k, :] = tot
    return out
