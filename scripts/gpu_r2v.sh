#!/bin/bash
# round 2, session v: team kernel with flush-and-continue list segments (16 warps for G1-43), ticket queue in the trajectory kernel
mkdir -p gpurun_out/r2v; O=gpurun_out/r2v
(timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
timeout 600 python scripts/bench_team.py g1_43_8192_esdf 512,1024,2048,4096,8192 2>&1 | tee $O/sweep.log
timeout 600 python scripts/bench_team.py g1_29_8192_esdf 512,1024,4096,8192 2>&1 | tee -a $O/sweep.log
for q in 0 1; do
  echo "CB200_QUEUE=$q"
  (CB200_QUEUE=$q timeout 900 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline) > $O/bench_q$q.log 2>&1
  tail -1 $O/bench_q$q.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms', d['ms_per_step'])
for k, v in d.get('other_workloads', {}).items(): print(' ', k, round(v.get('ms_per_step', -1), 4))
"
done
