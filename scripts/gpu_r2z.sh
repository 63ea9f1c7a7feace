#!/bin/bash
# round 2, session z: trajectory kernel synchronisation schedules A/B (neighbour flags with back-off / CTA barrier, double / single buffer)
mkdir -p gpurun_out/r2z; O=gpurun_out/r2z
(timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
for mode in "flags_double::" "flags_single:CB200_TRAJ_DOUBLE=0:" "cta_double:CB200_TRAJ_SYNC=1:" "cta_single:CB200_TRAJ_SYNC=1:CB200_TRAJ_DOUBLE=0"; do
  name=${mode%%:*}; rest=${mode#*:}; e1=${rest%%:*}; e2=${rest#*:}
  echo "== $name ($e1 $e2)"
  (env $e1 $e2 timeout 900 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads franka_trajopt_32x32_esdf_swept,franka_mpc_1024x30_esdf_swept,franka_mpc_knots_1024x30_esdf_swept,franka_mpc_knots_inkernel_1024x30_esdf_swept) > $O/bench_$name.log 2>&1
  tail -1 $O/bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k, v in d.get('other_workloads', {}).items(): print(' ', k, round(v.get('kernel_ms', -1), 4))
"
done
