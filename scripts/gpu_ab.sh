#!/bin/bash
# A/B of the 80-register arm variant + the two knots schedules
for mode in "CB200_ARM_REGCAP=0" "CB200_ARM_REGCAP=1"; do
  echo "$mode"
  for w in franka_ik_512x32_cuboid franka_16384_esdf; do
  env $mode timeout 200 python bench.py --workload $w --steps 300 --warmup 20 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   $w evals/s %.4g  ms %.4f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
  done
done
CB200_ARM_REGCAP=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ik-solve 1 --extra-workloads "" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('regcap0 ik_solve', d['ik_solve'])"
CB200_ARM_REGCAP=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ik-solve 1 --extra-workloads "franka_mpc_1024x30_esdf_swept,franka_mpc_knots_1024x30_esdf_swept,franka_mpc_knots_inkernel_1024x30_esdf_swept" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('regcap1 ik_solve', d['ik_solve']); print({k:(round(v.get('kernel_ms',0),4)) for k,v in d['other_workloads'].items()})"
timeout 300 python -m pytest tests/test_gpu_bspline.py tests/test_gpu_rollout.py -q -m gpu 2>&1 | tail -3
