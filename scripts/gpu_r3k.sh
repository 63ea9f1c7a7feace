#!/bin/bash
# round 2, session 3k: tile-scan loops unrolled by 2 in the big / team kernels
mkdir -p gpurun_out/r3k; O=gpurun_out/r3k
(timeout 900 python -m pytest tests/test_gpu_rollout.py -m gpu -q -x -p no:cacheprovider) 2>&1 | tail -2
timeout 600 python scripts/bench_team.py g1_29_8192_esdf 1024,8192 0,2 2>&1 | tee $O/sweep.log
timeout 600 python scripts/bench_team.py g1_43_8192_esdf 8192 0,2 2>&1 | tee -a $O/sweep.log
(timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads franka_16384_esdf,g1_29_8192_esdf,g1_43_8192_esdf) > $O/bench.log 2>&1
tail -1 $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'])
for k, v in d.get('other_workloads', {}).items(): print(' ', k, round(v.get('kernel_ms', -1), 4))
"
