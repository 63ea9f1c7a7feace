#!/bin/bash
mkdir -p gpurun_out/r10
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o gpurun_out/r10/prof_fused_g1 -f \
   python bench.py --workload g1_29_8192_esdf --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/r10/ncu_full_g1.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_traj -s 3 -c 1 -o gpurun_out/r10/prof_traj_mpc -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/r10/ncu_full_mpc.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o gpurun_out/r10/prof_fused_ik -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/r10/ncu_full_ik.log 2>&1
ls -la gpurun_out/r10
