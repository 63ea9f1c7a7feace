#!/bin/bash
# round 2, session h: ticket queue on the arm kernel, 18-warp humanoid build, restored CTA-phased dynamics kernel
mkdir -p gpurun_out/r2h; O=gpurun_out/r2h
(timeout 1200 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_zy_effort_cost.py tests/test_gpu_optim.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1
tail -3 $O/tests.log
B="python bench.py --steps 50 --warmup 5 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0"
W="franka_16384_esdf,g1_29_8192_esdf,g1_43_8192_esdf,franka_mpc_1024x30_esdf_swept_dynamics,franka_mpc_1024x30_esdf_swept_dynamics_host"
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['value_warm_l2']/1e8,3), {k: round(v.get('kernel_ms', -1), 4) for k, v in d['other_workloads'].items()})"; }
(timeout 300 $B --extra-workloads $W) > $O/default.log 2>&1; echo "default: $(show $O/default.log)"
(CB200_QUEUE=0 timeout 300 $B --extra-workloads $W) > $O/queue0.log 2>&1; echo "queue=0: $(show $O/queue0.log)"
(CB200_BIG=1 timeout 300 $B --extra-workloads $W) > $O/big1.log 2>&1; echo "big=1 (all robots): $(show $O/big1.log)"
(CB200_BIG=1 CB200_FORCE_NW=8 timeout 300 $B --extra-workloads franka_16384_esdf) > $O/big1_nw8.log 2>&1; echo "big=1 nw=8: $(show $O/big1_nw8.log)"
(CB200_BIG_MAXW=18 timeout 300 $B --extra-workloads g1_29_8192_esdf) > $O/maxw18.log 2>&1; echo "maxw=18: $(show $O/maxw18.log)"
(CB200_BIG_MAXW=18 CB200_FORCE_NW=17 timeout 300 $B --extra-workloads g1_29_8192_esdf) > $O/maxw18_17.log 2>&1; echo "maxw=18 nw=17: $(show $O/maxw18_17.log)"
(CB200_ARM_REGCAP=0 timeout 300 $B --extra-workloads "") > $O/regcap0.log 2>&1; echo "arm regcap=0: $(show $O/regcap0.log)"
(timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --edt 0 --rnea 0 --reference-design 0 --extra-workloads "") > $O/ik_solve.log 2>&1
tail -1 $O/ik_solve.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ik_solve', d['ik_solve']['solve_ms'], d['ik_solve']['success_rate'], 'sharded', {k: (round(v.get('rollout_ms_per_step',-1),4), round(v.get('solve_ms',-1),2), v.get('best_cost')) for k,v in d['sharded'].items()})"
