#!/bin/bash
# round 2, session a: baseline bench of the ABI-5 build + ncu captures that the kernel work of this round starts from
mkdir -p gpurun_out/r2a; O=gpurun_out/r2a
(timeout 900 python bench.py) > $O/bench.log 2>&1
(timeout 300 python scripts/bench_edt.py 128 256) > $O/edt_bench.jsonl 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ --launch-skip 6 -c 3 -o $O/prof_edt -f \
   python scripts/bench_edt.py 256 --no-ref) > $O/ncu_edt.log 2>&1
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused -s 3 -c 1 -o $O/prof_fused_g1 -f \
   python bench.py --workload g1_29_8192_esdf --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > $O/ncu_full_g1.log 2>&1
tail -1 $O/bench.log | cut -c1-6000; cat $O/edt_bench.jsonl | cut -c1-400
