#!/bin/bash
# round 2, session 3n: 80-register build of the trajectory kernel (24 instead of 16 warps per SM) on the MPC workloads
mkdir -p gpurun_out/r3n; O=gpurun_out/r3n
for t in 0 1; do
  (CB200_TRAJ80=$t timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads franka_trajopt_32x32_esdf_swept,franka_mpc_1024x30_esdf_swept,franka_mpc_knots_1024x30_esdf_swept) > $O/bench_t$t.log 2>&1
  tail -1 $O/bench_t$t.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k, v in d.get('other_workloads', {}).items(): print('traj80=$t', k, round(v.get('kernel_ms', -1), 4))
"
done
(CB200_TRAJ80=1 timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_bspline.py -m gpu -q -x -p no:cacheprovider -k "traj or mpc or knots") 2>&1 | tail -2
