#!/bin/bash
# round 2, session 3g: SMALL variant of the big kernel (Franka + ESDF): tests + bench
mkdir -p gpurun_out/r3g; O=gpurun_out/r3g
(timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
(timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads franka_16384_esdf,g1_29_8192_esdf) > $O/bench.log 2>&1
tail -1 $O/bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'])
for k, v in d.get('other_workloads', {}).items(): print(' ', k, round(v.get('kernel_ms', -1), 4))
"
