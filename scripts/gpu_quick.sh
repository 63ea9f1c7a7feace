#!/bin/bash
for mode in "CB200_LANE=1" "CB200_LANE=0"; do
  echo "$mode"
  for w in franka_ik_512x32_cuboid franka_16384_esdf; do
  env $mode timeout 200 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   $w evals/s %.4g  ms %.4f' % (d['value'], d['ms_per_step']))"
  done
done
