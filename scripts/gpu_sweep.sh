#!/bin/bash
# A/B sweep of the fused kernel's tuning knobs on the IK workload.
mkdir -p gpurun_out
for v in "" mb3 mb4; do for ps in 0 1 2; do
  echo "variant=$v phase_sync=$ps" 
  CB200_LIB_VARIANT=$v CB200_PHASE_SYNC=$ps timeout 200 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   evals/s %.4g  ms %.4f' % (d['value'], d['ms_per_step']))"
done; done
for v in "" mb3 mb4; do
  echo "variant=$v g1_29 / franka esdf / mpc"
  for w in g1_29_8192_esdf franka_16384_esdf franka_mpc_1024x30_esdf_swept; do
  CB200_LIB_VARIANT=$v timeout 200 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('   $w evals/s %.4g  ms %.4f' % (d['value'], d['ms_per_step']))"
  done
done
