#!/bin/bash
# round 2, session f: residency curve of the humanoid kernel (warps per CTA forced, 2 CTAs / SM)
mkdir -p gpurun_out/r2f; O=gpurun_out/r2f
for nw in 1 2 3 4 5; do
  (CB200_FORCE_NW=$nw timeout 300 python bench.py --workload g1_29_8192_esdf --steps 20 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/g1_nw$nw.log 2>&1
  echo "nw=$nw $(tail -1 $O/g1_nw$nw.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value_warm_l2'])")"
done
