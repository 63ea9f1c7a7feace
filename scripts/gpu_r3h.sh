#!/bin/bash
# round 2, session 3h: SMALL variant of the trajectory kernel: tests + MPC / trajopt bench
mkdir -p gpurun_out/r3h; O=gpurun_out/r3h
(timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
(timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads franka_trajopt_32x32_esdf_swept,franka_mpc_1024x30_esdf_swept,franka_mpc_knots_1024x30_esdf_swept,franka_mpc_knots_inkernel_1024x30_esdf_swept,franka_mpc_1024x30_esdf_swept_dynamics_host) > $O/bench.log 2>&1
tail -1 $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'])
for k, v in d.get('other_workloads', {}).items(): print(' ', k, round(v.get('kernel_ms', -1), 4))
"
