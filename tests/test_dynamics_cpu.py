"""RNEA inverse dynamics + adjoint on the CPU: the numpy oracle against finite differences and physics identities, and
the product's __host__ __device__ arithmetic (curobo_b200/csrc/cb200_dynamics.cuh, host-compiled) against the oracle."""
import ctypes as C

import numpy as np
import pytest

from dynamics_cases import CASES, make_case, model_args, pack_cache
from helpers import hostmath, ptr
from oracle import dynamics_oracle as do


def hm_forward(c):
    B, nl, D = c["B"], c["nl"], c["D"]
    tau, cache = np.zeros((B, D), np.float32), np.zeros((B, nl * 20), np.float32)
    m = [np.ascontiguousarray(x) for x in model_args(c)]
    hostmath().hm_rnea_forward(ptr(tau), ptr(cache), ptr(c["q"]), ptr(c["qd"]), ptr(c["qdd"]), *[ptr(x) for x in m],
                               ptr(c["starts"]), ptr(c["order"]), C.c_int(B), C.c_int(nl), C.c_int(D), C.c_int(c["n_levels"]))
    return tau, cache


def hm_backward(c, cache):
    B, nl, D = c["B"], c["nl"], c["D"]
    gq, gqd, gqdd = (np.zeros((B, D), np.float32) for _ in range(3))
    m = [np.ascontiguousarray(x) for x in model_args(c)]
    hostmath().hm_rnea_backward(ptr(gq), ptr(gqd), ptr(gqdd), ptr(c["grad_tau"]), ptr(c["q"]), ptr(c["qd"]), ptr(cache),
                                *[ptr(x) for x in m], ptr(c["starts"]), ptr(c["order"]), C.c_int(B), C.c_int(nl), C.c_int(D),
                                C.c_int(c["n_levels"]))
    return gq, gqd, gqdd


def test_oracle_physics_identities():
    c = make_case("franka", 6, 11)
    m = model_args(c)
    z = np.zeros_like(c["q"])
    t_g = do.rnea_forward(c["q"], z, z, *m)[0]                       # gravity torque only
    t_0 = do.rnea_forward(c["q"], z, z, *m[:-1], np.zeros(6, np.float32))[0]
    assert np.abs(t_0).max() < 1e-6 and np.abs(t_g).max() > 1.0       # no gravity, at rest: zero torque
    t1 = do.rnea_forward(c["q"], z, c["qdd"], *m)[0]
    t2 = do.rnea_forward(c["q"], z, 2 * c["qdd"], *m)[0]
    assert np.abs((t2 - t_g) - 2 * (t1 - t_g)).max() < 2e-4 * np.abs(t2).max()    # M(q) qdd is linear in qdd
    # mass matrix from unit accelerations is symmetric
    D = c["D"]
    M = np.stack([do.rnea_forward(c["q"][:1], z[:1], np.eye(D, dtype=np.float32)[i:i + 1], *m)[0][0] - t_g[0] for i in range(D)])
    assert np.abs(M - M.T).max() < 2e-4 * np.abs(M).max()
    assert np.all(np.linalg.eigvalsh(0.5 * (M + M.T).astype(np.float64)) > 0)


@pytest.mark.parametrize("robot,B,seed", [("franka", 3, 1), ("g1_29", 2, 2)])
def test_oracle_backward_matches_finite_differences(robot, B, seed):
    """The reference's hand-written adjoint (restated) == d <grad_tau, tau> / d (q, qd, qdd) of its forward.
    (For robots with prismatic joints the reference's motion_cross_S puts the Coriolis term of those joints in the
    angular slots; the adjoint was derived for the generic operator, so only revolute trees are checked for q / qd.)"""
    c = make_case(robot, B, seed)
    m = model_args(c)
    tau, cache = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    gq, gqd, gqdd = do.rnea_backward(c["grad_tau"], c["q"], c["qd"], cache, *m)
    loss = lambda q, qd, qdd: (do.rnea_forward(q, qd, qdd, *m)[0].astype(np.float64) * c["grad_tau"]).sum(1)  # noqa: E731
    eps = 1e-2
    rng = np.random.default_rng(0)
    checks = (("qdd", gqdd),) if robot != "franka" else (("q", gq), ("qd", gqd), ("qdd", gqdd))
    for name, g in checks:
        for d in rng.choice(c["D"], size=min(c["D"], 5), replace=False):
            args = {"q": c["q"].copy(), "qd": c["qd"].copy(), "qdd": c["qdd"].copy()}
            hi, lo = {k: v.copy() for k, v in args.items()}, {k: v.copy() for k, v in args.items()}
            hi[name][:, d] += eps
            lo[name][:, d] -= eps
            num = (loss(hi["q"], hi["qd"], hi["qdd"]) - loss(lo["q"], lo["qd"], lo["qdd"])) / (2 * eps)
            assert np.allclose(num, g[:, d], rtol=2e-2, atol=2e-3 * np.abs(g).max()), (name, d)


@pytest.mark.parametrize("robot,B,seed", CASES)
def test_host_math_vs_oracle(robot, B, seed):
    c = make_case(robot, B, seed)
    m = model_args(c)
    tau_w, cache_w = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    tau, cache = hm_forward(c)
    sc = np.abs(tau_w).max()
    assert np.allclose(tau, tau_w, rtol=1e-4, atol=2e-5 * sc)
    cw = pack_cache(cache_w, c["nl"])
    assert np.allclose(cache, cw, rtol=1e-4, atol=2e-5 * np.abs(cw).max())
    want = do.rnea_backward(c["grad_tau"], c["q"], c["qd"], cache_w, *m)
    got = hm_backward(c, cache)
    for g, w in zip(got, want):
        assert np.allclose(g, w, rtol=2e-4, atol=5e-5 * np.abs(w).max())


def test_effort_cost_composition_gradient_by_finite_differences():
    """tau = RNEA(q, qd, qdd) -> effort channel of the STATE cost -> RNEA adjoint: the composed gradient is the derivative
    of the summed cost (float64 central differences along random directions of q, qd, qdd)."""
    from dynamics_cases import effort_cost_oracle, effort_cost_setup
    c, shape, jerk, dt, limits, weight, act, reg = effort_cost_setup(B=3, H=4)
    cost, g, tau = effort_cost_oracle(c, shape, jerk, dt, limits, weight, act, reg)
    assert (np.abs(g[0]).sum() > 0) and np.isfinite(cost).all()
    # the effort hinge is active somewhere and the energy term contributes
    assert ((tau < limits["tau"][0] + act[4]) | (tau > limits["tau"][1] - act[4])).mean() > 0.1
    rng = np.random.default_rng(0)
    for key, gi in (("q", 0), ("qd", 1), ("qdd", 2)):
        d = rng.normal(size=c[key].shape)
        d /= np.linalg.norm(d)
        eps = 1.6e-2  # the oracle is float32: smaller steps drown in the rounding of the summed cost
        vals = []
        for sgn in (+1, -1):
            c2 = dict(c)
            c2[key] = (c[key].astype(np.float64) + sgn * eps * d).astype(np.float32)
            vals.append(float(effort_cost_oracle(c2, shape, jerk, dt, limits, weight, act, reg)[0].astype(np.float64).sum()))
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float((g[gi].reshape(d.shape).astype(np.float64) * d).sum())
        assert abs(fd - an) <= 1.5e-2 * max(abs(an), abs(fd), 1.0), (key, fd, an)


@pytest.mark.parametrize("nl,B,seed,mimic", __import__("dynamics_cases").RANDOM_TREES)
def test_host_math_vs_oracle_on_random_trees(nl, B, seed, mimic):
    """Synthetic trees with every joint type, fixed links, negative multipliers, offsets and mimic joints (links sharing a
    joint index): the product's row functions against the oracle, forward and adjoint."""
    from dynamics_cases import random_tree_case
    c = random_tree_case(nl, B, seed, mimic)
    m = model_args(c)
    tau_w, cache_w = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    tau, cache = hm_forward(c)
    assert np.allclose(tau, tau_w, rtol=1e-4, atol=3e-5 * np.abs(tau_w).max())
    cw = pack_cache(cache_w, c["nl"])
    assert np.allclose(cache, cw, rtol=1e-4, atol=3e-5 * np.abs(cw).max())
    want = do.rnea_backward(c["grad_tau"], c["q"], c["qd"], cache_w, *m)
    got = hm_backward(c, cache)
    for g, w in zip(got, want):
        assert np.allclose(g, w, rtol=3e-4, atol=1e-4 * max(np.abs(w).max(), 1e-6))


def test_oracle_qdd_adjoint_on_a_mimic_tree_by_finite_differences():
    """tau is linear in qdd (tau = M(q) qdd + ...), so d<grad_tau, tau>/d qdd is exact by differencing -- with mimic joints
    the same column of M collects several links."""
    from dynamics_cases import random_tree_case
    c = random_tree_case(40, 2, 33, True)
    assert int((c["rm"].joint_map >= 0).sum()) > c["D"], "the case must contain mimic links"
    m = model_args(c)
    tau, cache = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    gqdd = do.rnea_backward(c["grad_tau"], c["q"], c["qd"], cache, *m)[2]
    for d in range(0, c["D"], 3):
        e = np.zeros_like(c["qdd"])
        e[:, d] = 1.0
        t1 = do.rnea_forward(c["q"], c["qd"], c["qdd"] + e, *m)[0].astype(np.float64)
        num = ((t1 - tau.astype(np.float64)) * c["grad_tau"]).sum(1)
        assert np.allclose(num, gqdd[:, d], rtol=5e-3, atol=5e-3 * np.abs(gqdd).max()), d
