#!/bin/bash
# round 2, session 3c: where is the instruction-cache cliff?  IK kernel with parts of the row switched off at run time
mkdir -p gpurun_out/r3c; O=gpurun_out/r3c
for tag in full no_self no_scene no_pose no_self_scene; do
  (timeout 300 ncu --metrics smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio --clock-control none -k regex:rollout_fused -s 20 -c 1 --csv python scripts/ablate_icache.py $tag) > $O/$tag.log 2>&1
  echo "== $tag"; grep -E "no_instruction|inst_executed.sum|time_duration|issue_active.avg|: .* us" $O/$tag.log | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
done
