"""Seeded inputs for the RNEA tests: real kinematic trees (Franka chain, G1 humanoid with a 3-prismatic + 3-revolute
floating base and mimic-free joints) with random, physically plausible inertial parameters."""
import numpy as np

from curobo_b200.robot_model import load_robot
from oracle import dynamics_oracle as do


def make_case(robot, B, seed):
    rm = load_robot(robot)
    nl, D = rm.num_links, rm.num_dof
    rng = np.random.default_rng(seed)
    mc = np.zeros((nl, 4), np.float32)
    mc[:, :3] = rng.uniform(-0.05, 0.05, (nl, 3))
    mc[:, 3] = rng.uniform(0.2, 3.0, nl)
    inn = np.zeros((nl, 8), np.float32)
    for k in range(nl):
        A = rng.normal(size=(3, 3))
        Im = (A @ A.T) * 0.004 + np.eye(3) * 0.002
        inn[k, :6] = [Im[0, 0], Im[1, 1], Im[2, 2], Im[0, 1], Im[0, 2], Im[1, 2]]
    lo, hi = rm.position_limits[0], rm.position_limits[1]
    q = rng.uniform(np.maximum(lo, -2.0), np.minimum(hi, 2.0), (B, D)).astype(np.float32)
    qd = rng.uniform(-1.5, 1.5, (B, D)).astype(np.float32)
    qdd = rng.uniform(-3.0, 3.0, (B, D)).astype(np.float32)
    starts, order = do.tree_levels(rm.link_map)
    return dict(rm=rm, B=B, nl=nl, D=D, mc=mc, inn=inn, q=q, qd=qd, qdd=qdd, grad_tau=rng.normal(size=(B, D)).astype(np.float32),
                gravity=np.array([0, 0, 0, 0, 0, 9.81], np.float32), starts=starts, order=order, n_levels=len(starts) - 1)


def model_args(c):
    rm = c["rm"]
    return (rm.fixed_transforms.astype(np.float32), c["mc"], c["inn"], rm.joint_map_type.astype(np.int8),
            rm.joint_map.astype(np.int16), rm.link_map.astype(np.int16), rm.joint_offset_map.astype(np.float32), c["gravity"])


def pack_cache(cache, nl):
    """oracle cache dict -> the kernels' [B, nl*20] layout (v, a, f, pad2)."""
    B = cache["v"].shape[0]
    out = np.zeros((B, nl, 20), np.float32)
    out[:, :, 0:6], out[:, :, 6:12], out[:, :, 12:18] = cache["v"], cache["a"], cache["f"]
    return out.reshape(B, nl * 20)


CASES = [("franka", 9, 1), ("g1_29", 5, 2), ("g1_43", 3, 3)]


def effort_cost_oracle(c, shape, jerk, dt, limits, weight, act, reg):
    """Dynamics-aware STATE cost composed from the two oracles: tau = RNEA(q, qd, qdd) feeds the effort channel of
    cspace_state_cost; grad_tau walks back through the RNEA adjoint (robot_state_transition.py:380-389 +
    wp_cspace_state.py:21-285).  Returns cost [B,H,D], (g_p, g_v, g_a, g_j), tau."""
    from oracle import rollout_oracle as O
    B, H, D = shape
    m = model_args(c)
    tau, cache = do.rnea_forward(c["q"], c["qd"], c["qdd"], *m)
    r = lambda x: np.asarray(x, np.float32).reshape(B, H, D)  # noqa: E731
    cost, g = O.cspace_state_cost(r(c["q"]), r(c["qd"]), r(c["qdd"]), jerk, dt, limits, weight, act, reg, True, True,
                                  effort=r(tau))
    gq, gqd, gqdd = do.rnea_backward(g[4].reshape(B * H, D), c["q"], c["qd"], cache, *m)
    return cost, (g[0] + r(gq), g[1] + r(gqd), g[2] + r(gqdd), g[3]), r(tau)


def effort_cost_setup(robot="franka", B=6, H=5, seed=21):
    c = make_case(robot, B * H, seed)
    rm = c["rm"]
    rng = np.random.default_rng(seed + 1)
    D = c["D"]
    jerk = rng.normal(0, 50.0, size=(B, H, D)).astype(np.float32)
    dt = rng.uniform(0.02, 0.1, size=B).astype(np.float32)
    tau0 = do.rnea_forward(c["q"], c["qd"], c["qdd"], *model_args(c))[0]
    # effort limits placed inside the range of the torques of this case so that the hinge is active on a good fraction
    elim = np.stack([np.quantile(tau0, 0.2, axis=0), np.quantile(tau0, 0.8, axis=0)]).astype(np.float32)
    limits = dict(p=rm.position_limits, v=rm.velocity_limits, a=rm.acceleration_limits, j=rm.jerk_limits, tau=elim)
    weight = np.array([5000.0, 500.0, 50.0, 5.0, 20.0], np.float32)
    act = np.array([0.01, 0.01, 0.01, 0.01, 0.5], np.float32)
    reg = np.array([10.0, 1.0, 0.01, 0.05, 0.3], np.float32)
    return c, (B, H, D), jerk, dt, limits, weight, act, reg


class _Tree:
    """The RobotModel fields the RNEA cases read, for a synthetic tree."""


def random_tree_case(nl, B, seed, mimic=True):
    """A random kinematic tree: random parents (parents precede children), every joint type of the reference (-1 fixed, 0-2
    prismatic x/y/z, 3-5 revolute x/y/z), random fixed transforms, random (multiplier, offset) per joint incl. negative
    multipliers, and -- with `mimic` -- links that share a joint index (mimic joints: tau and the gradients then accumulate
    over links, robot/parser/parser_urdf.py:220-224)."""
    rng = np.random.default_rng(seed)
    link_map = np.zeros(nl, np.int16)
    jtype = np.full(nl, -1, np.int8)
    jmap = np.full(nl, -1, np.int16)
    joff = np.zeros((nl, 2), np.float32)
    joff[:, 0] = 1.0
    ft = np.zeros((nl, 3, 4), np.float32)
    D = 0
    for k in range(nl):
        A = rng.normal(size=(3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        ft[k, :, :3] = Q
        ft[k, :, 3] = rng.uniform(-0.3, 0.3, 3)
        if k == 0:
            link_map[0] = -1
            continue
        link_map[k] = rng.integers(max(0, k - 4), k)
        if rng.random() < 0.2:
            continue                                     # fixed link
        jtype[k] = rng.integers(0, 6)
        if mimic and D > 0 and rng.random() < 0.25:
            jmap[k] = rng.integers(0, D)                 # shares an existing joint
        else:
            jmap[k] = D
            D += 1
        joff[k] = [rng.choice([-1.0, 1.0]) * rng.uniform(0.5, 1.5), rng.uniform(-0.3, 0.3)]
    D = max(D, 1)
    rm = _Tree()
    rm.num_links, rm.num_dof = nl, D
    rm.fixed_transforms, rm.joint_map_type, rm.joint_map, rm.link_map, rm.joint_offset_map = ft, jtype, jmap, link_map, joff
    mc = np.zeros((nl, 4), np.float32)
    mc[:, :3] = rng.uniform(-0.05, 0.05, (nl, 3))
    mc[:, 3] = rng.uniform(0.2, 3.0, nl)
    inn = np.zeros((nl, 8), np.float32)
    for k in range(nl):
        A = rng.normal(size=(3, 3))
        Im = (A @ A.T) * 0.004 + np.eye(3) * 0.002
        inn[k, :6] = [Im[0, 0], Im[1, 1], Im[2, 2], Im[0, 1], Im[0, 2], Im[1, 2]]
    q = rng.uniform(-1.5, 1.5, (B, D)).astype(np.float32)
    qd = rng.uniform(-1.5, 1.5, (B, D)).astype(np.float32)
    qdd = rng.uniform(-3.0, 3.0, (B, D)).astype(np.float32)
    lm = link_map.copy()
    lm[0] = 0                                            # tree_levels reads the parent of the root as itself
    starts, order = do.tree_levels(lm)
    return dict(rm=rm, B=B, nl=nl, D=D, mc=mc, inn=inn, q=q, qd=qd, qdd=qdd, grad_tau=rng.normal(size=(B, D)).astype(np.float32),
                gravity=np.array([0, 0, 0, 0, 0, 9.81], np.float32), starts=starts, order=order, n_levels=len(starts) - 1)


RANDOM_TREES = [(5, 4, 31, True), (17, 3, 32, True), (40, 2, 33, True), (64, 2, 34, False), (9, 5, 35, False)]
