#!/bin/bash
# round 2, session 3f: instruction-fetch stall of the other kernels (Franka + ESDF through the big kernel, G1-29, MPC)
mkdir -p gpurun_out/r3f; O=gpurun_out/r3f
for wl in franka_16384_esdf g1_29_8192_esdf g1_43_8192_esdf franka_mpc_1024x30_esdf_swept; do
  (timeout 300 ncu --metrics smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:rollout_ -s 3 -c 1 --csv python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/$wl.log 2>&1
  echo "== $wl"; grep -E "stalled|inst_executed.sum|time_duration|issue_active.avg" $O/$wl.log | python -c "
import sys,csv
for l in sys.stdin:
    r=next(csv.reader([l])); print('  ', r[4][22:60], r[-3].replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), r[-1])
"
done
