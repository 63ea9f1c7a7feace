"""Seeded inputs for the optimizer-kernel tests."""
import numpy as np


def lbfgs_case(seed, B, V, m, fill=None, negative_curvature=False, zero_history=False):
    """Random but consistent L-BFGS state: `fill` previous (s, y) pairs with y = A s for an SPD A (positive curvature),
    plus the current point / gradient."""
    rng = np.random.default_rng(seed)
    fill = m if fill is None else fill
    A = rng.normal(size=(B, V, V)).astype(np.float32)
    A = (A @ A.transpose(0, 2, 1) + 0.5 * np.eye(V, dtype=np.float32)).astype(np.float32)
    S = np.zeros((m, B, V), np.float32)
    Y = np.zeros((m, B, V), np.float32)
    rho = np.zeros((m, B), np.float32)
    if not zero_history:
        for i in range(m - fill, m):
            s = (rng.normal(size=(B, V)) * 0.1).astype(np.float32)
            y = np.einsum("bij,bj->bi", A, s).astype(np.float32)
            S[i], Y[i] = s, y
            rho[i] = 1.0 / np.sum(s * y, -1)
    x0 = rng.normal(size=(B, V)).astype(np.float32)
    g0 = rng.normal(size=(B, V)).astype(np.float32)
    s = (rng.normal(size=(B, V)) * 0.1).astype(np.float32)
    q = (x0 + s).astype(np.float32)
    y = np.einsum("bij,bj->bi", A, s).astype(np.float32)
    if negative_curvature:
        y[::2] = -y[::2]          # y.s < 0 for every other problem -> rho = 0 in stable mode
    gq = (g0 + y).astype(np.float32)
    return dict(rho=rho, Y=Y, S=S, q=q, grad_q=gq, x_0=x0, grad_0=g0)


LBFGS_CASES = [
    dict(seed=1, B=37, V=7, m=7),                       # IK shape (lbfgs_ik.yml history 7)
    dict(seed=2, B=5, V=4, m=3),
    dict(seed=3, B=9, V=16, m=7, fill=3),
    dict(seed=4, B=6, V=29, m=15),
    dict(seed=5, B=4, V=32, m=5),
    dict(seed=6, B=3, V=112, m=27),                     # 16 knots x 7 dof, trajopt history 27
    dict(seed=7, B=3, V=168, m=15),
    dict(seed=8, B=2, V=1000, m=31),
    dict(seed=9, B=33, V=7, m=7, negative_curvature=True),
    dict(seed=10, B=8, V=7, m=7, zero_history=True),    # first call after reset
]


def lbfgs_id(kw):
    return f"B{kw['B']}-V{kw['V']}-m{kw['m']}" + ("-neg" if kw.get("negative_curvature") else "") + \
        ("-fresh" if kw.get("zero_history") else "")


def line_search_case(seed, B, V, n, mode="mixed"):
    rng = np.random.default_rng(seed)
    mags = np.array([0.0, 0.1, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0][:n], np.float32) if n <= 8 else np.linspace(0, 2, n).astype(np.float32)
    x = rng.normal(size=(B, V)).astype(np.float32)
    p = rng.normal(size=(B, V)).astype(np.float32)
    A = rng.normal(size=(B, V, V)).astype(np.float32) / np.sqrt(V)
    A = (A @ A.transpose(0, 2, 1) + 0.1 * np.eye(V, dtype=np.float32)).astype(np.float32)
    bvec = rng.normal(size=(B, V)).astype(np.float32)
    g_at = lambda z: (np.einsum("bij,bj->bi", A, z) - bvec).astype(np.float32)  # noqa: E731
    f_at = lambda z: (0.5 * np.einsum("bi,bij,bj->b", z, A, z) - np.sum(bvec * z, -1)).astype(np.float32)  # noqa: E731
    if mode != "ascent":
        p = (-g_at(x) * rng.uniform(0.05, 3.0, size=(B, 1))).astype(np.float32)   # descent directions of varied length
    xs = (x[:, None, :] + mags[None, :, None] * p[:, None, :]).astype(np.float32)
    cost = np.stack([f_at(xs[:, i]) for i in range(n)], 1)
    grad = np.stack([g_at(xs[:, i]) for i in range(n)], 1)
    best_cost = (cost[:, 0] + rng.normal(size=B).astype(np.float32) * 0.5).astype(np.float32)
    return dict(search_cost=cost, search_action=xs, search_gradient=grad, step_direction=p, magnitudes=mags,
                best_cost=best_cost, best_action=rng.normal(size=(B, V)).astype(np.float32),
                best_iteration=rng.integers(0, 5, B).astype(np.int16), current_iteration=rng.integers(5, 20, B).astype(np.int16))


LS_CASES = [
    dict(seed=1, B=41, V=7, n=4),
    dict(seed=2, B=7, V=4, n=4),
    dict(seed=3, B=5, V=16, n=8),
    dict(seed=4, B=6, V=32, n=4),
    dict(seed=5, B=3, V=112, n=4),
    dict(seed=6, B=3, V=200, n=7),
    dict(seed=7, B=9, V=7, n=4, mode="ascent"),
    dict(seed=8, B=4, V=3, n=6),                        # more candidates than lanes in the group (G = 4)
]


def ls_id(kw):
    return f"B{kw['B']}-V{kw['V']}-n{kw['n']}" + ("-" + kw["mode"] if "mode" in kw else "")
