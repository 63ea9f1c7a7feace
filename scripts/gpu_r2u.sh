#!/bin/bash
# round 2, session u: compute-sanitizer (memcheck, racecheck, synccheck) on the team kernel tests; ncu capture of the team kernel
mkdir -p gpurun_out/r2u; O=gpurun_out/r2u
for tool in memcheck racecheck synccheck; do
  (timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "team_kernel" -p no:cacheprovider) > $O/sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/sanitizer_$tool.log | tail -3
done
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:team_kernel -s 20 -c 1 -o $O/team2_g1_29_1024 -f python scripts/bench_team.py g1_29_8192_esdf 1024 2) > $O/ncu_team.log 2>&1; tail -2 $O/ncu_team.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:team_kernel -s 20 -c 1 -o $O/team2_g1_43_8192 -f python scripts/bench_team.py g1_43_8192_esdf 8192 2) > $O/ncu_team43.log 2>&1; tail -2 $O/ncu_team43.log
ls -la $O
