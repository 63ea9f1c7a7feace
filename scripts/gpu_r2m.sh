#!/bin/bash
# round 2, session m: cull pre-pass of the humanoid kernel (four look-ups in flight): parity + A/B
mkdir -p gpurun_out/r2m; O=gpurun_out/r2m
(timeout 1500 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1; tail -3 $O/tests.log
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0"
W="franka_16384_esdf,g1_29_8192_esdf,g1_43_8192_esdf"
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), {k: round(v.get('kernel_ms', -1), 4) for k, v in d['other_workloads'].items()})"; }
for i in 1 2; do (timeout 300 $B --extra-workloads $W) > $O/run$i.log 2>&1; echo "run $i: $(show $O/run$i.log)"; done
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused_big -s 3 -c 1 -o $O/prof_big_g1 -f \
   $B --workload g1_29_8192_esdf --extra-workloads "") > $O/ncu_big.log 2>&1
