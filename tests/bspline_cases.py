"""Shared seeded inputs for the B-spline tests (CPU host-math tests and GPU parity tests)."""
import numpy as np

from oracle import bspline_oracle as bo


def make_case(seed, B, nk, D, steps, degree, implicit, n_start=2, n_goal=3, mixed_implicit=False, zero_goal_rates=False):
    rng = np.random.default_rng(seed)
    T = bo.padded_horizon_for(nk, degree, steps)
    c = dict(B=B, nk=nk, D=D, steps=steps, degree=degree, T=T)
    c["knots"] = rng.normal(size=(B, nk, D)).astype(np.float32)
    c["start"] = tuple((rng.normal(size=(n_start, D)) * s).astype(np.float32) for s in (1.0, 0.5, 0.3, 0.2))
    gs = (1.0, 0.0, 0.0, 0.0) if zero_goal_rates else (1.0, 0.5, 0.3, 0.2)
    c["goal"] = tuple((rng.normal(size=(n_goal, D)) * s).astype(np.float32) for s in gs)
    c["start_idx"] = rng.integers(0, n_start, B).astype(np.int32)
    c["goal_idx"] = rng.integers(0, n_goal, B).astype(np.int32)
    c["traj_dt"] = rng.uniform(0.01, 0.2, n_goal).astype(np.float32)
    imp = np.full(n_goal, int(implicit), np.uint8)
    if mixed_implicit:
        imp = (np.arange(n_goal) % 2).astype(np.uint8)
    c["implicit"] = imp
    c["grads"] = tuple(rng.normal(size=(B, T, D)).astype(np.float32) for _ in range(4))
    return c


CASES = [
    dict(seed=1, B=5, nk=8, D=7, steps=4, degree=3, implicit=False),
    dict(seed=2, B=5, nk=8, D=7, steps=4, degree=4, implicit=False),
    dict(seed=3, B=5, nk=8, D=7, steps=4, degree=5, implicit=False),
    dict(seed=4, B=4, nk=16, D=7, steps=2, degree=4, implicit=True),
    dict(seed=5, B=4, nk=10, D=9, steps=1, degree=3, implicit=True),
    dict(seed=6, B=3, nk=6, D=35, steps=8, degree=5, implicit=True),
    dict(seed=7, B=6, nk=12, D=7, steps=4, degree=4, implicit=False, mixed_implicit=True),
    dict(seed=8, B=3, nk=7, D=5, steps=3, degree=4, implicit=False),   # non power-of-two steps
    dict(seed=9, B=2, nk=6, D=3, steps=16, degree=3, implicit=True),
]


def case_id(kw):
    return f"deg{kw['degree']}-nk{kw['nk']}-s{kw['steps']}-{'imp' if kw['implicit'] else 'rep'}" + \
        ("-mixed" if kw.get("mixed_implicit") else "")
