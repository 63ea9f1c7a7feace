"""Optimizer-kernel oracle on the CPU: pinned against golden vectors produced by the REFERENCE's own torch
implementation (tests/golden/make_optim_golden.py), plus properties of the L-BFGS step and the Wolfe selection."""
import os

import numpy as np
import pytest

from optim_cases import LBFGS_CASES, LS_CASES, lbfgs_case, lbfgs_id, line_search_case, ls_id
from oracle import optim_oracle as oo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lbfgs_reference_torch.npz")


@pytest.mark.parametrize("tag", ["ik", "trajopt", "small"])
def test_lbfgs_oracle_vs_reference_torch_golden(tag):
    """6 consecutive iterations of the reference's jit_lbfgs_update_buffers + jit_lbfgs_compute_step_direction:
    history buffers bit-equal, rho and the step to 2e-6 relative (torch sums with matmul, the kernel with a tree)."""
    g = np.load(GOLD)
    for it in range(6):
        k = lambda n: g[f"{tag}_{n}"][it]  # noqa: E731
        step, rho, Y, S, x0, g0 = oo.lbfgs_step(k("rho_in"), k("y_in"), k("s_in"), k("q"), k("grad_q"), k("x0_in"),
                                                k("g0_in"), 0.01, True)
        assert np.array_equal(Y, k("y_out")) and np.array_equal(S, k("s_out"))
        assert np.allclose(rho, k("rho_out"), rtol=2e-6, atol=0)
        assert np.allclose(step, k("step"), rtol=0, atol=2e-6 * np.abs(k("step")).max())
        assert np.array_equal(x0, k("q")) and np.array_equal(g0, k("grad_q"))


def test_tree_sum_matches_exact_sum_and_order():
    rng = np.random.default_rng(0)
    for n in (1, 3, 7, 31, 32, 33, 112, 1000):
        v = rng.normal(size=(5, n)).astype(np.float32)
        assert np.allclose(oo.tree_sum(v), v.astype(np.float64).sum(-1), rtol=1e-5, atol=1e-5)
    # association for 7 elements: ((v0+v4)+(v2+v6)) + ((v1+v5)+v3)
    v = np.array([[1e8, 1.0, -1e8, 1.0, 1.0, 1.0, 1.0]], np.float32)
    want = np.float32(np.float32(np.float32(v[0, 0] + v[0, 4]) + np.float32(v[0, 2] + v[0, 6]))
                      + np.float32(np.float32(v[0, 1] + v[0, 5]) + v[0, 3]))
    assert oo.tree_sum(v)[0] == want


@pytest.mark.parametrize("kw", LBFGS_CASES, ids=lbfgs_id)
def test_lbfgs_step_properties(kw):
    c = lbfgs_case(**kw)
    step, rho, Y, S, x0, g0 = oo.lbfgs_step(c["rho"], c["Y"], c["S"], c["q"], c["grad_q"], c["x_0"], c["grad_0"], 0.01, True)
    assert np.isfinite(step).all()
    # history rolled left, newest pair appended
    assert np.array_equal(Y[:-1], c["Y"][1:]) and np.array_equal(S[:-1], c["S"][1:])
    assert np.array_equal(S[-1], (c["q"] - c["x_0"]).astype(np.float32))
    assert np.array_equal(Y[-1], (c["grad_q"] - c["grad_0"]).astype(np.float32))
    if kw.get("negative_curvature"):
        assert np.all(rho[-1, ::2] == 0) and np.all(rho[-1, 1::2] > 0)
    elif not kw.get("zero_history"):
        # positive-curvature history -> the two-loop matrix is positive definite -> descent direction
        assert np.all(np.sum(step * c["grad_q"], -1) < 0)
    # secant equation of the newest pair: H y = s  =>  step(g) - step(g - y) = -s ... checked via linearity in g
    c2 = dict(c)
    g_alt = (c["grad_q"] * 1.0).astype(np.float32)
    step2 = oo.lbfgs_step(c["rho"], c["Y"], c["S"], c["q"], g_alt, c["x_0"], c["grad_0"], 0.01, True)[0]
    assert np.array_equal(step, step2)  # deterministic


@pytest.mark.parametrize("kw", LS_CASES, ids=ls_id)
def test_line_search_oracle_properties(kw):
    c = line_search_case(**kw)
    for strong, approx in ((False, True), (False, False), (True, False)):
        o = oo.line_search(c["best_cost"], c["best_action"], c["best_iteration"], c["current_iteration"], 10, 0.0, 0.0,
                           c["search_cost"], c["search_action"], c["search_gradient"], c["step_direction"], c["magnitudes"],
                           1e-5, 0.9, strong, approx)
        B, n = c["search_cost"].shape
        sel, ex = o["selected_idx"][:, 0], o["exploration_idx"][:, 0]
        assert np.all((sel >= 0) & (sel < n))
        # selected is an Armijo point whenever any candidate satisfies Armijo (candidate 0 always does: c <= c0)
        g0 = np.sum(c["search_gradient"][:, 0] * c["step_direction"], -1)
        arm = c["search_cost"] <= c["search_cost"][:, :1] + 1e-5 * c["magnitudes"][None] * g0[:, None] + 1e-6
        assert np.all(arm[np.arange(B), sel])
        if approx and not strong:
            assert np.all(ex[sel == 0] == 1) and np.all(ex[sel != 0] == sel[sel != 0])
        else:
            assert np.array_equal(ex, sel)
        rows = np.arange(B)
        assert np.array_equal(o["selected_action"], c["search_action"][rows, sel])
        assert np.array_equal(o["selected_gradient"], c["search_gradient"][rows, sel])
        assert np.array_equal(o["current_iteration"], c["current_iteration"] + 1)
        delta = c["best_cost"] - o["selected_cost"]
        upd = (delta > 0) & (delta / (c["best_cost"] + np.float32(1e-6)) > 0)   # relative test: line_search_helpers.cuh:33-35
        assert np.array_equal(o["best_cost"][upd], o["selected_cost"][upd])
        assert np.array_equal(o["best_cost"][~upd], c["best_cost"][~upd])
        assert np.array_equal(o["best_action"][upd], o["selected_action"][upd])


def test_line_search_picks_largest_admissible_step():
    """1-D convex quadratic f(x) = x^2 / 2 from x = 1 along p = -1: Armijo holds for alpha in (0, 2(1-c1)); curvature
    (weak Wolfe, c2 = 0.9) needs alpha >= 0.1.  With candidates 0, 0.1, 0.5, 1.0 the largest passing both is 1.0."""
    mags = np.array([0.0, 0.1, 0.5, 1.0], np.float32)
    xs = (1.0 - mags)[None, :, None].astype(np.float32)
    cost = (0.5 * xs[..., 0] ** 2).astype(np.float32)
    grad = xs.copy()
    o = oo.line_search(np.array([10.0], np.float32), np.zeros((1, 1), np.float32), np.zeros(1, np.int16),
                       np.zeros(1, np.int16), 10, 0.0, 0.0, cost, xs, grad, np.array([[-1.0]], np.float32), mags, 1e-5, 0.9,
                       False, True)
    assert o["selected_idx"][0, 0] == 3 and o["selected_cost"][0] == 0.0
    # along an ascent direction nothing but alpha = 0 is admissible: selected 0, exploration 1 (approx Wolfe)
    xs2 = (1.0 + mags)[None, :, None].astype(np.float32)
    o = oo.line_search(np.array([10.0], np.float32), np.zeros((1, 1), np.float32), np.zeros(1, np.int16),
                       np.zeros(1, np.int16), 10, 0.0, 0.0, (0.5 * xs2[..., 0] ** 2).astype(np.float32), xs2, xs2.copy(),
                       np.array([[1.0]], np.float32), mags, 1e-5, 0.9, False, True)
    assert o["selected_idx"][0, 0] == 0 and o["exploration_idx"][0, 0] == 1
