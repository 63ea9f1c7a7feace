#!/bin/bash
# round 2, session k: ncu --set full of the IK and MPC kernels (current build) for per-line attribution; mesh bench after the box early-out
mkdir -p gpurun_out/r2k; O=gpurun_out/r2k
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads \"\""
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused_kernel -s 3 -c 1 -o $O/prof_ik -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_ik.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_traj_kernel -s 3 -c 1 -o $O/prof_mpc -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_mpc.log 2>&1
(timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv \
   python bench.py --steps 5 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_launches.log 2>&1
(timeout 300 python scripts/bench_mesh.py) > $O/mesh_bench.json 2>&1; cat $O/mesh_bench.json | cut -c1-600
ls -la $O
