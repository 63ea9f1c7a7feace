#!/bin/bash
# round 2, session l: identity-rotation fast path A/B (all BASELINE configs), parity suite on the new build
mkdir -p gpurun_out/r2l; O=gpurun_out/r2l
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x) > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0"
W="franka_16384_esdf,franka_trajopt_32x32_esdf_swept,franka_mpc_1024x30_esdf_swept,g1_29_8192_esdf,g1_43_8192_esdf"
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value_warm_l2']/1e8,3), {k: round(v.get('kernel_ms', -1), 4) for k, v in d['other_workloads'].items()})"; }
for i in 1 2; do (timeout 300 $B --extra-workloads $W) > $O/run$i.log 2>&1; echo "run $i: $(show $O/run$i.log)"; done
