#!/bin/bash
# One GPU session: smoke, full GPU test suites (all three schedules), bench, ncu launch list + full captures.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
(timeout 300 python __graft_entry__.py --smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
(timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider) > gpurun_out/gpu_tests.log 2>&1
(CB200_TILE=1 timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q -p no:cacheprovider -k "not traj") > gpurun_out/rollout_tile.log 2>&1
(CB200_LANE=1 timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q -p no:cacheprovider -k "not traj") > gpurun_out/rollout_lane.log 2>&1
(timeout 600 python bench.py) > gpurun_out/bench.log 2>&1
(timeout 600 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/bench_reference.log 2>&1
if [ "$1" != "noprof" ]; then
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/ncu_launches.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o gpurun_out/prof_fused_ik -f \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/ncu_full.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o gpurun_out/prof_fused_g1 -f \
   python bench.py --workload g1_29_8192_esdf --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/ncu_full_g1.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_traj -s 3 -c 1 -o gpurun_out/prof_traj_mpc -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/ncu_full_mpc.log 2>&1
fi
tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/gpu_tests.log; tail -2 gpurun_out/rollout_tile.log; tail -2 gpurun_out/rollout_lane.log
tail -1 gpurun_out/bench.log | cut -c1-3000; tail -1 gpurun_out/bench_reference.log | cut -c1-400
