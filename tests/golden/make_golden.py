#!/usr/bin/env python
"""Generate the committed golden fixtures tests/golden/*.npz from the oracle (seeded inputs).

    python tests/golden/make_golden.py        # rewrites the .npz files

The reference itself cannot run here (no warp / cuda.core / GPU), so the fixtures are
oracle outputs; the oracle is pinned separately (tests/test_oracle_golden.py and, on the GPU box,
tests/test_gpu_vs_reference_kernels.py against the reference's compiled CUDA kernels)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from curobo_b200.robot_model import load_robot  # noqa: E402
from curobo_b200.rollout import RolloutConfig  # noqa: E402
from curobo_b200.world import make_benchmark_cuboid_world  # noqa: E402
from helpers import random_q, random_walk_q, small_voxel_world  # noqa: E402
from oracle import rollout_oracle as O  # noqa: E402


def goal_from_q(rm, qg):
    _, _, p, qt = O.fk_forward(rm, qg)
    return p[:, :, None, :].copy(), qt[:, :, None, :].copy()      # [G,L,1,3/4]


def build_cases():
    out = {}
    # config 1: Franka FK + self-collision, batch 64 (BASELINE.json configs[0])
    rm = load_robot("franka")
    q = random_q(rm, 64, seed=11)
    cum, sph, pos, quat = O.fk_forward(rm, q)
    c, g, k = O.self_collision(sph, rm.sphere_padding, rm.collision_pairs, 5000.0)
    out["franka_fk_self_b64"] = dict(q=q, cumul=cum, spheres=sph, link_pos=pos, link_quat=quat, self_cost=c,
                                     self_grad=g, self_pair=k)
    # IK-style fused rollout: cuboid world, pose goal, c-space position (small batch)
    cfg = RolloutConfig.ik()
    qb = random_q(rm, 48, seed=12)[:, None, :]
    gp, gq = goal_from_q(rm, random_q(rm, 4, seed=13))
    idx = (np.arange(48) % 4).astype(np.int32)
    r = O.rollout_cost_grad(rm, qb, cfg.to_oracle_cfg(1), world_cuboid=make_benchmark_cuboid_world(),
                            goal_pos=gp, goal_quat=gq, idxs_goal=idx)
    out["franka_ik_rollout_b48"] = dict(q=qb, goal_pos=gp, goal_quat=gq, idxs_goal=idx, cost_bh=r["cost_bh"],
                                        grad_q=r["grad_q"], self_cost=r["self_cost"], scene_cost=r["scene_cost"],
                                        pose_cost=r["pose_cost"], cspace_cost=r["cspace_cost"])
    # ESDF discrete rollout on the humanoid (small batch)
    rg = load_robot("g1_29")
    qh = random_q(rg, 6, seed=14, scale=0.6)[:, None, :]
    cfgh = RolloutConfig(self_weight=5000.0, scene_weight=5000.0, scene_activation=0.02, cspace_type="position",
                         cspace_weight=(5000.0, 0, 0, 0, 0), cspace_activation=(0.01, 0, 0, 0, 0))
    vox = small_voxel_world()
    r = O.rollout_cost_grad(rg, qh, cfgh.to_oracle_cfg(14), world_voxel=vox)
    out["g1_29_esdf_rollout_b6"] = dict(q=qh, cost_bh=r["cost_bh"], grad_q=r["grad_q"], self_cost=r["self_cost"],
                                        scene_cost=r["scene_cost"], cspace_cost=r["cspace_cost"])
    # swept + speed metric scene collision on a Franka trajectory batch
    qt = random_walk_q(rm, 4, 12, seed=15)
    _, spt, _, _ = O.fk_forward(rm, qt.reshape(-1, 7))
    spt = spt.reshape(4, 12, -1, 4)
    c, g = O.scene_collision(spt, 100000.0, 0.0025, make_benchmark_cuboid_world(), vox, sweep=True, speed_dt=0.05)
    out["franka_swept_b4_h12"] = dict(q=qt, spheres=spt, scene_cost=c, scene_grad=g)
    return out


if __name__ == "__main__":
    for name, arrays in build_cases().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})
