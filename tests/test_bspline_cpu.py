"""B-spline knot -> state and adjoint on the CPU: the numpy oracle's algebraic properties, and the product's
__host__ __device__ arithmetic (curobo_b200/csrc/cb200_bspline.cuh, host-compiled in tests/hostmath) against it."""
import ctypes as C

import numpy as np
import pytest

from oracle import bspline_oracle as bo
from bspline_cases import CASES, case_id, make_case
from helpers import hostmath, ptr


def hm_forward(lib, c, interp_h=None, interp_dt=None):
    B, T, D, nk = c["B"], c["T"], c["D"], c["nk"]
    outs = [np.zeros((B, T, D), np.float32) for _ in range(4)]
    dt = c["traj_dt"] if interp_h is None else np.array([interp_dt], np.float32)
    ih = None if interp_h is None else np.ascontiguousarray(interp_h, np.int32)
    rc = lib.hm_bspline_forward(*[ptr(o) for o in outs], ptr(c["knots"]), *[ptr(x) for x in c["start"]],
                                *[ptr(x) for x in c["goal"]], ptr(c["start_idx"]), ptr(c["goal_idx"]), ptr(dt),
                                ptr(c["implicit"]), ptr(ih), C.c_int(B), C.c_int(T), C.c_int(D), C.c_int(nk),
                                C.c_int(c["degree"]))
    assert rc == 0
    return outs


def hm_backward(lib, c):
    B, T, D, nk = c["B"], c["T"], c["D"], c["nk"]
    out = np.zeros((B, nk, D), np.float32)
    rc = lib.hm_bspline_backward(ptr(out), *[ptr(g) for g in c["grads"]], ptr(c["traj_dt"]), ptr(c["goal_idx"]),
                                 ptr(c["implicit"]), C.c_int(B), C.c_int(T), C.c_int(D), C.c_int(nk), C.c_int(c["degree"]))
    assert rc == 0
    return out


def oracle_forward(c, **kw):
    return bo.bspline_forward(c["knots"], c["start"], c["goal"], c["start_idx"], c["goal_idx"], c["traj_dt"], c["implicit"],
                              c["T"], c["degree"], **kw)


def state_tol(c, k):
    """Derivative k is divided by knot_dt^k: absolute error scales with the magnitude of the output."""
    return 2e-5


@pytest.mark.parametrize("degree", [3, 4, 5])
def test_basis_partition_of_unity(degree):
    t = np.linspace(0, 1, 33).astype(np.float32)
    bp, bv, ba, bj = bo.basis_rows(degree, t)
    assert np.allclose(bp.sum(-1), 1.0, atol=2e-6)
    for b in (bv, ba, bj):
        assert np.allclose(b.sum(-1), 0.0, atol=3e-5)
    # derivative consistency: d/dt position basis == velocity basis (finite differences in float64)
    h = 1e-3
    tm = np.float32(0.4)
    b0 = bo.basis_rows(degree, np.asarray(tm - h, np.float32))[0].astype(np.float64)
    b1 = bo.basis_rows(degree, np.asarray(tm + h, np.float32))[0].astype(np.float64)
    assert np.allclose((b1 - b0) / (np.float32(tm + h).astype(np.float64) - np.float32(tm - h)), bo.basis_rows(degree, tm)[1],
                       atol=2e-3)


@pytest.mark.parametrize("kw", CASES, ids=case_id)
def test_oracle_boundary_states(kw):
    """Row 0 reproduces the start state (as far as the degree can: cubic has no jerk control,
    bspline_boundary_constraint.cuh:39); in replicate mode the padded last row rests at the last knot."""
    c = make_case(**kw)
    p, v, a, j, dt = oracle_forward(c)
    s = [x[c["start_idx"]] for x in c["start"]]
    kdt = (np.maximum(c["traj_dt"][c["goal_idx"]], 1e-6) * c["steps"])[:, None]
    assert np.allclose(p[:, 0], s[0], atol=1e-5)
    assert np.allclose(v[:, 0] * kdt, s[1] * kdt, atol=2e-5)
    assert np.allclose(a[:, 0] * kdt**2, s[2] * kdt**2, atol=5e-5)
    if c["degree"] >= 4:
        assert np.allclose(j[:, 0] * kdt**3, s[3] * kdt**3, atol=2e-4)
    rep = c["implicit"][c["goal_idx"]] == 0
    assert np.allclose(p[rep, -1], c["knots"][rep, -1], atol=1e-5)
    assert np.allclose((v[:, -1] * kdt)[rep], 0, atol=2e-5)
    assert np.array_equal(dt, c["traj_dt"][c["goal_idx"]])


def test_oracle_implicit_goal_reached_one_segment_before_the_end():
    """Implicit mode re-uses the start-state coefficient table for the goal (bspline_boundary_constraint.cuh:319-352):
    the goal state is met at the start of the LAST segment; with zero goal rates the spline then rests there."""
    kw = dict(seed=11, B=4, nk=9, D=7, steps=4, degree=4, implicit=True, zero_goal_rates=True)
    c = make_case(**kw)
    p, v, a, j, _ = oracle_forward(c)
    h_goal = (bo.total_knots(c["nk"], c["degree"]) - 1) * c["steps"]
    g = c["goal"][0][c["goal_idx"]]
    assert np.allclose(p[:, h_goal], g, atol=1e-5)
    assert np.allclose(p[:, -1], g, atol=1e-5)
    assert np.abs(v[:, h_goal:]).max() < 1e-3


@pytest.mark.parametrize("kw", CASES, ids=case_id)
def test_oracle_backward_is_the_adjoint(kw):
    """<J du, g> == <du, J^T g>: the reference's hand-written backward is the exact adjoint of its forward
    (including the replicate-tail accumulation and the zeroed last knot in implicit mode)."""
    c = make_case(**kw)
    rng = np.random.default_rng(99)
    du = rng.normal(size=c["knots"].shape).astype(np.float32)
    base = oracle_forward(c)[:4]
    c2 = dict(c, knots=(c["knots"] + du).astype(np.float32))
    pert = oracle_forward(c2)[:4]
    lhs = sum(((x2.astype(np.float64) - x1) * g).sum() for x1, x2, g in zip(base, pert, c["grads"]))
    gk = bo.bspline_backward(*c["grads"], c["traj_dt"], c["goal_idx"], c["implicit"], c["nk"], c["degree"])
    rhs = (gk.astype(np.float64) * du).sum()
    scale = sum(np.abs((x2.astype(np.float64) - x1) * g).sum() for x1, x2, g in zip(base, pert, c["grads"]))
    assert abs(lhs - rhs) < 2e-5 * scale


@pytest.mark.parametrize("kw", CASES, ids=case_id)
def test_host_math_forward_vs_oracle(kw):
    c = make_case(**kw)
    lib = hostmath()
    got = hm_forward(lib, c)
    want = oracle_forward(c)[:4]
    for k, (g, w) in enumerate(zip(got, want)):
        assert np.allclose(g, w, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(w).max())), f"derivative {k}"


@pytest.mark.parametrize("kw", CASES, ids=case_id)
def test_host_math_backward_vs_oracle(kw):
    c = make_case(**kw)
    got = hm_backward(hostmath(), c)
    want = bo.bspline_backward(*c["grads"], c["traj_dt"], c["goal_idx"], c["implicit"], c["nk"], c["degree"])
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5 * np.abs(want).max())
    dead = c["implicit"][c["goal_idx"]] != 0
    assert np.all(got[dead, -1] == 0)


def test_host_math_single_dt_vs_oracle():
    c = make_case(seed=21, B=6, nk=8, D=7, steps=4, degree=4, implicit=False)
    T = 70
    c = dict(c, T=T)
    interp_h = np.array([52, 39, 69, 13, 200, 26], np.int32)
    got = hm_forward(hostmath(), c, interp_h=interp_h, interp_dt=0.025)
    want = bo.bspline_forward(c["knots"], c["start"], c["goal"], c["start_idx"], c["goal_idx"], c["traj_dt"], c["implicit"],
                              T, c["degree"], interpolation_horizon=interp_h, interpolation_dt=np.float32(0.025))[:4]
    for g, w in zip(got, want):
        assert np.allclose(g, w, rtol=2e-5, atol=2e-5 * max(1.0, np.abs(w).max()))
    # rows past a trajectory's own horizon hold its end state
    assert np.allclose(got[0][1, 40:], got[0][1, 39], atol=1e-6)
