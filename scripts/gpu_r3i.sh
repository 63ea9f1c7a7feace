#!/bin/bash
# round 2, session 3i: final build -- full GPU suite, smoke, bench (both arms), launch list + dram traffic of the headline kernel
mkdir -p gpurun_out/r3i; O=gpurun_out/r3i
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; tail -1 $O/smoke.log
(timeout 900 python bench.py) > $O/bench.log 2>&1; tail -1 $O/bench.log > $O/bench_line.json
(timeout 600 python bench.py --impl reference) > $O/bench_ref.log 2>&1; tail -1 $O/bench_ref.log > $O/bench_reference_line.json; cut -c1-300 $O/bench_reference_line.json
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/launches.log 2>&1
(timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:rollout_ -s 10 -c 1 --csv --log-file $O/traffic_ik.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/traffic.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/r3i/bench_line.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "kernel", d["roofline"]["kernel"], "frac", d["roofline"]["frac"])
for k, v in d.get("other_workloads", {}).items():
    print(" ", k, round(v.get("kernel_ms", -1), 4), v.get("kernel"))
print("sharded", {k: (round(v.get("rollout_ms_per_step", -1), 4), round(v.get("solve_ms", -1), 2)) if "skipped" not in v else "skipped" for k, v in d["sharded"].items()})
print("ik_solve", d.get("ik_solve", {}).get("solve_ms")); print("edt", d.get("edt")); print("clocks", d.get("clocks"))
print("refdesign", {k: round(v["speedup"], 2) for k, v in d.get("reference_design_gpu", {}).get("workloads", {}).items()})
print("cpu_baseline", d.get("cpu_baseline"))
PY
grep -c rollout $O/launches.csv; grep -E "dram__bytes|time_duration" $O/traffic_ik.csv | cut -d, -f5,13- | head -5
