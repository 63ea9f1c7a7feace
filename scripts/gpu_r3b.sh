#!/bin/bash
# round 2, session 3b: IK kernel, CTA lockstep experiment (instruction fetch sharing)
mkdir -p gpurun_out/r3b; O=gpurun_out/r3b
for ls in 0 1 0 1; do
  (CB200_LOCKSTEP=$ls timeout 600 python bench.py --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads "") > $O/bench_l$ls.log 2>&1
  tail -1 $O/bench_l$ls.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('lockstep $ls: value', d['value'], 'ms', d['ms_per_step'], 'warm', d.get('value_warm_l2'))
"
done
(CB200_LOCKSTEP=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o $O/prof_ik_lockstep -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_ik.log 2>&1
tail -1 $O/ncu_ik.log
