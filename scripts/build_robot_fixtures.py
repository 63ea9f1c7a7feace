#!/usr/bin/env python
"""Generate curobo_b200/content/robots/*.npz from the reference's own robot content.

Run in the build container (needs /root/reference; the GPU box only reads the committed .npz):
    python scripts/build_robot_fixtures.py
Inputs: /root/reference/curobo/content/configs/robot/{franka,unitree_g1_29dof_retarget,unitree_g1}.yml
        and the URDFs they name under content/assets/.
"""
import os
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from curobo_b200.robot_model import build_robot_model  # noqa: E402

REF = os.environ.get("CUROBO_REFERENCE", "/root/reference")
CONTENT = os.path.join(REF, "curobo", "content")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "curobo_b200", "content", "robots")

ROBOTS = {"franka": "franka.yml", "g1_29": "unitree_g1_29dof_retarget.yml", "g1_43": "unitree_g1.yml"}


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, yml in ROBOTS.items():
        d = yaml.safe_load(open(os.path.join(CONTENT, "configs", "robot", yml)))
        k = d.get("robot_cfg", d)["kinematics"]
        urdf = os.path.join(CONTENT, "assets", k["urdf_path"])
        m = build_robot_model(name, urdf, k)
        m.save(os.path.join(OUT, name + ".npz"))
        print(f"{name}: links={m.num_links} dof={m.num_dof} spheres={m.num_spheres} "
              f"tool_frames={m.num_tool_frames} pairs={m.collision_pairs.shape[0]} "
              f"blocks_per_batch={m.num_blocks_per_batch}")


if __name__ == "__main__":
    main()
