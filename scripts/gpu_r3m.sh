#!/bin/bash
# round 2, session 3m: final build (arms + ESDF routed to the 80-register build at full batch): full suite, smoke, bench both arms
mkdir -p gpurun_out/r3m; O=gpurun_out/r3m
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; tail -1 $O/smoke.log
(timeout 900 python bench.py) > $O/bench.log 2>&1; tail -1 $O/bench.log > $O/bench_line.json
(timeout 600 python bench.py --impl reference) > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log > $O/bench_reference_line.json; cut -c1-200 $O/bench_reference_line.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3m/bench_line.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel", d["roofline"]["kernel"], "frac", d["roofline"]["frac"])
for k, v in d.get("other_workloads", {}).items():
    print(" ", k, round(v.get("kernel_ms", -1), 4), v.get("kernel"))
print("sharded", {k: (round(v.get("rollout_ms_per_step", -1), 4), round(v.get("solve_ms", -1), 2)) if "skipped" not in v else "skipped" for k, v in d["sharded"].items()})
print("ik_solve", d.get("ik_solve", {}).get("solve_ms")); print("clocks", d.get("clocks"))
print("refdesign", {k: (round(v["fused_ms"], 4), round(v["speedup"], 2)) for k, v in d.get("reference_design_gpu", {}).get("workloads", {}).items()})
PY
