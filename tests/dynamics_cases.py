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
