#!/bin/bash
# One GPU session: tests, bench, register-cap sweep, ncu launch list + full capture of the fused kernel.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt
(timeout 300 python __graft_entry__.py --smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --maxfail=40 -p no:cacheprovider) > gpurun_out/parity.log 2>&1
(timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q --maxfail=40 -p no:cacheprovider) > gpurun_out/rollout.log 2>&1
(timeout 400 python bench.py --steps 200 --warmup 10) > gpurun_out/bench.log 2>&1
if [ "$1" != "noprof" ]; then
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extra-workloads "") > gpurun_out/ncu_launches.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 2 -o gpurun_out/prof_fused_ik -f \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extra-workloads "") > gpurun_out/ncu_full.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o gpurun_out/prof_fused_g1 -f \
   python bench.py --workload g1_29_8192_esdf --steps 3 --warmup 3 --no-cpu-baseline --extra-workloads "") > gpurun_out/ncu_full_g1.log 2>&1
fi
tail -3 gpurun_out/smoke.log; tail -15 gpurun_out/parity.log; tail -15 gpurun_out/rollout.log
tail -1 gpurun_out/bench.log | cut -c1-2500
