#!/bin/bash
# round 2, session 3a: fresh ncu capture of the IK kernel (instruction-cache footprint analysis)
mkdir -p gpurun_out/r3a; O=gpurun_out/r3a
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o $O/prof_ik -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_ik.log 2>&1
tail -2 $O/ncu_ik.log
