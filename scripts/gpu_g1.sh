#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -3
for w in franka_ik_512x32_cuboid g1_29_8192_esdf g1_43_8192_esdf franka_mpc_1024x30_esdf_swept; do
timeout 200 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$w evals/s %.4g  ms %.4f' % (d['value'], d['ms_per_step']))"
done
