#!/bin/bash
# round 2, session t: full GPU suite + bench line with the team kernels selected by default
mkdir -p gpurun_out/r2t; O=gpurun_out/r2t
(timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
(timeout 900 python bench.py) > $O/bench.log 2>&1; tail -1 $O/bench.log > $O/bench_line.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2t/bench_line.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k, v in d.get("other_workloads", {}).items():
    print(" ", k, {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("ms_per_step", "value")})
print("sharded", {k: (round(v.get("rollout_ms_per_step", -1), 4), round(v.get("solve_ms", -1), 2)) if "skipped" not in v else "skipped" for k, v in d["sharded"].items()})
print("ik_solve", d.get("ik_solve")); print("clocks", d.get("clocks"))
PY
