#!/bin/bash
# round 2, session x: ncu capture of the trajectory kernel on the MPC workload (stall breakdown, per-line samples)
mkdir -p gpurun_out/r2x; O=gpurun_out/r2x
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_traj -s 3 -c 1 -o $O/prof_traj_mpc -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_full_mpc.log 2>&1
tail -3 $O/ncu_full_mpc.log; ls -la $O
