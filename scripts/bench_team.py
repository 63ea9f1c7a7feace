"""Small-batch sweep of the humanoid rollout: rows x team size (CB200_TEAM = 0: one warp per row, the big-robot kernel).
Timing: CUDA events around 100 launches after warm-up, L2 flushed between launches is NOT done (small batches: the working set is
the 32 MiB ESDF + KB of rows, resident in L2 in every arm) -- this is an A/B of kernels, not a bench line."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402

names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["g1_29_8192_esdf"]
rows = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "128,256,512,1024,1536,2048,4096".split(","))]
for name in names:
    full = bench.make_workload(name)
    for n in rows:
        wl = bench.shard_workload(full, 0, n)
        line = f"{name} rows {n:5d}:"
        ref = None
        for team in (sys.argv[3].split(",") if len(sys.argv) > 3 else ("0", "2", "4", "")):
            team = "" if team == "auto" else team
            if team:
                os.environ["CB200_TEAM"] = team
            else:
                os.environ.pop("CB200_TEAM", None)
            eng = bench.build_engine(wl, "cuda:0")
            q = torch.as_tensor(wl["q"]).cuda()
            for _ in range(10):
                out = eng.evaluate_action(q)
            torch.cuda.synchronize()
            g = out.grad_q.clone()
            if ref is None:
                ref = g
            err = float((g - ref).abs().max() / ref.abs().max())
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100):
                eng.evaluate_action(q)
            b.record()
            torch.cuda.synchronize()
            line += f"  team={team or 'auto'} {a.elapsed_time(b) / 100 * 1e3:7.1f} us (rel {err:.1e})"
        print(line, flush=True)
