#!/bin/bash
# round 2, session g: warp-specialised dynamics kernel + big-robot (humanoid) kernel on the GPU: parity, sanitizer, A/B timings
mkdir -p gpurun_out/r2g; O=gpurun_out/r2g
(timeout 1200 python -m pytest tests/test_gpu_rollout.py tests/test_gpu_zz_edt.py tests/test_gpu_zy_effort_cost.py tests/test_gpu_sharded_solve.py tests/test_gpu_reference_callsites.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1
tail -4 $O/tests.log
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware") > $O/dyn_memcheck.log 2>&1
(timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware_rollout") > $O/dyn_racecheck.log 2>&1
(timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rollout.py -m gpu -q -p no:cacheprovider -k "big_robot") > $O/big_memcheck.log 2>&1
(timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_rollout.py -m gpu -q -p no:cacheprovider -k "big_robot and g1_29") > $O/big_racecheck.log 2>&1
tail -2 $O/dyn_memcheck.log; tail -2 $O/dyn_racecheck.log; tail -2 $O/big_memcheck.log; tail -2 $O/big_racecheck.log
(timeout 300 python scripts/bench_edt.py 128 256) > $O/edt_bench.jsonl 2>&1; cut -c1-300 $O/edt_bench.jsonl
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0"
DYNW="franka_mpc_1024x30_esdf_swept,franka_mpc_1024x30_esdf_swept_dynamics_host,franka_mpc_1024x30_esdf_swept_dynamics,franka_mpc_knots_1024x30_esdf_swept_dynamics"
for rw in 8 7 6 5; do
  (CB200_DYN_ROW_WARPS=$rw timeout 300 $B --extra-workloads $DYNW) > $O/bench_dyn_rw$rw.log 2>&1
  echo "dyn row warps $rw: $(tail -1 $O/bench_dyn_rw$rw.log | python -c "import json,sys; d=json.loads(sys.stdin.read())['other_workloads']; print({k[-24:]: round(v.get('kernel_ms', -1), 4) for k, v in d.items()})")"
done
G1W="g1_29_8192_esdf,g1_43_8192_esdf,franka_16384_esdf"
(CB200_BIG=0 timeout 300 $B --extra-workloads $G1W) > $O/bench_big0.log 2>&1
echo "big=0: $(tail -1 $O/bench_big0.log | python -c "import json,sys; d=json.loads(sys.stdin.read())['other_workloads']; print({k: round(v.get('kernel_ms', -1), 4) for k, v in d.items()})")"
for nw in 16 12 10 8; do
  (CB200_BIG=1 CB200_FORCE_NW=$nw timeout 300 $B --extra-workloads $G1W) > $O/bench_big1_nw$nw.log 2>&1
  echo "big=1 nw=$nw: $(tail -1 $O/bench_big1_nw$nw.log | python -c "import json,sys; d=json.loads(sys.stdin.read())['other_workloads']; print({k: round(v.get('kernel_ms', -1), 4) for k, v in d.items()})")"
done
(timeout 300 $B --extra-workloads $G1W) > $O/bench_default.log 2>&1
echo "default: $(tail -1 $O/bench_default.log | python -c "import json,sys; d=json.loads(sys.stdin.read())['other_workloads']; print({k: round(v.get('kernel_ms', -1), 4) for k, v in d.items()})")"
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fused_big -s 3 -c 1 -o $O/prof_big_g1 -f \
   $B --workload g1_29_8192_esdf --extra-workloads "") > $O/ncu_big.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_traj_dyn -s 3 -c 1 -o $O/prof_traj_dyn -f \
   $B --workload franka_mpc_1024x30_esdf_swept_dynamics --extra-workloads "") > $O/ncu_dyn.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ --launch-skip 6 -c 3 -o $O/prof_edt -f \
   python scripts/bench_edt.py 256 --no-ref) > $O/ncu_edt.log 2>&1
