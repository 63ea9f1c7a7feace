#!/bin/bash
# round 2, session 3e: pair kernel (rows in pairs, phase by phase): tests, ncu stall metrics, bench A/B
mkdir -p gpurun_out/r3e; O=gpurun_out/r3e
(timeout 900 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider) 2>&1 | tail -2
(timeout 300 ncu --metrics smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__warps_active.avg.per_cycle_active --clock-control none -k regex:rollout_fused -s 20 -c 1 --csv python scripts/ablate_icache.py full) > $O/full.log 2>&1
grep -E "no_instruction|inst_executed.sum|time_duration|issue_active.avg|registers|warps_active|Kernel Name" $O/full.log | python -c "
import sys,csv
for l in sys.stdin:
    r=next(csv.reader([l])); print('  ', r[4][:60] if len(r)>4 else '', r[-3], r[-1])
"
for p in 1 0 1 0; do
(CB200_PAIR=$p timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads "") > $O/bench_p$p.log 2>&1
tail -1 $O/bench_p$p.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pair $p: value', d['value'], 'ms', d['ms_per_step'], 'warm', d.get('value_warm_l2'), 'e2e', d['e2e']['value'])
"
done
