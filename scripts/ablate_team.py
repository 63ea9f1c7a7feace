"""Where a humanoid row's latency goes at small batch: phase ablation (weights off) x team size, few rows (one row per SM)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from curobo_b200.rollout import RolloutConfig  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "g1_29_8192_esdf"
for n in (int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "128,1024").split(",")):
    for tag, sw, cw in (("full", 5000.0, 5000.0), ("no_scene", 5000.0, 0.0), ("no_self", 0.0, 5000.0), ("fk_bwd_only", 0.0, 0.0)):
        line = f"{name} rows {n:5d} {tag:12s}"
        for team in ("0", "2", "4"):
            os.environ["CB200_TEAM"] = team
            wl = bench.shard_workload(bench.make_workload(name), 0, n)
            c = wl["cfg"]
            wl["cfg"] = RolloutConfig(**{**c.__dict__, "self_weight": sw, "scene_weight": cw})
            eng = bench.build_engine(wl, "cuda:0")
            q = torch.as_tensor(wl["q"]).cuda()
            for _ in range(10):
                eng.evaluate_action(q)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100):
                eng.evaluate_action(q)
            b.record()
            torch.cuda.synchronize()
            line += f"  team={team} {a.elapsed_time(b) / 100 * 1e3:7.1f} us"
        print(line, flush=True)
