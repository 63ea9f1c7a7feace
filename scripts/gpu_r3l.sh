#!/bin/bash
# round 2, session 3l: Franka + ESDF: big kernel (SMALL) vs the 80-register arm build of the standard kernel
mkdir -p gpurun_out/r3l; O=gpurun_out/r3l
for mode in "default::" "arm80:CB200_BIG=0:CB200_ARM_ESDF=1" "std128:CB200_BIG=0:"; do
  name=${mode%%:*}; rest=${mode#*:}; e1=${rest%%:*}; e2=${rest#*:}
  (env $e1 $e2 timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads franka_16384_esdf) > $O/bench_$name.log 2>&1
  tail -1 $O/bench_$name.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k, v in d.get('other_workloads', {}).items(): print('$name', k, round(v.get('kernel_ms', -1), 4), v.get('kernel'))
"
done
