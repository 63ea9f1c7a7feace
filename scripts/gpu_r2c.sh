#!/bin/bash
# round 2, session c: the chunked dynamics kernel on the GPU (parity, memcheck, timing)
mkdir -p gpurun_out/r2c; O=gpurun_out/r2c
(timeout 600 python -m pytest tests/test_gpu_zy_effort_cost.py tests/test_gpu_dynamics.py tests/test_gpu_zx_dynamics_trees.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware") > $O/memcheck.log 2>&1
(timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware_rollout") > $O/racecheck.log 2>&1
(timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads franka_mpc_1024x30_esdf_swept,franka_mpc_1024x30_esdf_swept_dynamics_host,franka_mpc_1024x30_esdf_swept_dynamics,franka_mpc_knots_1024x30_esdf_swept_dynamics) > $O/bench.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_traj_dyn -s 3 -c 1 -o $O/prof_traj_dyn -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept_dynamics --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu.log 2>&1
tail -4 $O/tests.log; tail -4 $O/memcheck.log; tail -4 $O/racecheck.log; tail -1 $O/bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(json.dumps(d['other_workloads'], indent=0))"
