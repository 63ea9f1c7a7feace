import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from curobo_b200.rollout import RolloutConfig
name = sys.argv[1] if len(sys.argv) > 1 else "g1_29_8192_esdf"
for tag, sw, cw in (("full", 5000.0, 5000.0), ("no_scene", 5000.0, 0.0), ("no_self", 0.0, 5000.0), ("fk_bwd_only", 0.0, 0.0)):
    wl = bench.make_workload(name)
    c = wl["cfg"]
    wl["cfg"] = RolloutConfig(**{**c.__dict__, "self_weight": sw, "scene_weight": cw})
    eng = bench.build_engine(wl, "cuda:0")
    q = torch.as_tensor(wl["q"]).cuda()
    for _ in range(5): eng.evaluate_action(q)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): eng.evaluate_action(q)
    b.record(); torch.cuda.synchronize()
    print(f"{name} {tag:12s} {a.elapsed_time(b)/50*1e3:8.1f} us (warm L2)")
