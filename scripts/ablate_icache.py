"""One launch of the IK workload with parts of the row switched off at run time (weights = 0: the code stays in the binary but is not
fetched), for ncu: how does the "no instruction" stall react to a smaller hot code footprint?  usage: ablate_icache.py <tag>"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from curobo_b200.rollout import RolloutConfig  # noqa: E402

tag = sys.argv[1]
wl = bench.make_workload("franka_ik_512x32_cuboid")
c = wl["cfg"]
over = {"full": {}, "no_self": {"self_weight": 0.0}, "no_scene": {"scene_weight": 0.0}, "no_pose": {"pose_weight": (0.0, 0.0)},
        "no_self_scene": {"self_weight": 0.0, "scene_weight": 0.0}}[tag]
wl["cfg"] = RolloutConfig(**{**c.__dict__, **over})
if tag == "no_pose":
    wl["goal"] = None
eng = bench.build_engine(wl, "cuda:0")
q = torch.as_tensor(wl["q"]).cuda()
for _ in range(5):
    eng.evaluate_action(q)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    eng.evaluate_action(q)
b.record()
torch.cuda.synchronize()
print(f"{tag}: {a.elapsed_time(b) / 50 * 1e3:.1f} us (warm L2)")
