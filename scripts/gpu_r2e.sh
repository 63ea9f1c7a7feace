#!/bin/bash
# round 2, session e: banded EDT + ESDF builder, the reference's call sites over the b200 backend, the warp-specialised dynamics
# kernel (parity, sanitizer, row-warp sweep), residency curve of the humanoid kernel
mkdir -p gpurun_out/r2e; O=gpurun_out/r2e
(timeout 900 python -m pytest tests/test_gpu_zz_edt.py tests/test_gpu_reference_callsites.py tests/test_gpu_zy_effort_cost.py tests/test_gpu_sharded_solve.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1
(timeout 300 compute-sanitizer --tool memcheck python scripts/bench_edt.py 64 --no-ref) > $O/memcheck.log 2>&1
(timeout 300 compute-sanitizer --tool racecheck python scripts/bench_edt.py 64 --no-ref) > $O/racecheck.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware") > $O/dyn_memcheck.log 2>&1
(timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware_rollout") > $O/dyn_racecheck.log 2>&1
(timeout 300 python scripts/bench_edt.py 128 256) > $O/edt_bench.jsonl 2>&1
DYNW="franka_mpc_1024x30_esdf_swept,franka_mpc_1024x30_esdf_swept_dynamics_host,franka_mpc_1024x30_esdf_swept_dynamics,franka_mpc_knots_1024x30_esdf_swept_dynamics"
for rw in 8 7 6 5; do
  (CB200_DYN_ROW_WARPS=$rw timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads $DYNW) > $O/bench_dyn_rw$rw.log 2>&1
  echo "row warps $rw: $(tail -1 $O/bench_dyn_rw$rw.log | python -c "import json,sys; d=json.loads(sys.stdin.read())['other_workloads']; print({k[-22:]: round(v.get('kernel_ms', -1), 4) for k, v in d.items()})")"
done
for nw in 1 2 3 4 5; do
  (CB200_FORCE_NW=$nw timeout 300 python bench.py --workload g1_29_8192_esdf --steps 20 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/g1_nw$nw.log 2>&1
  echo "g1 nw=$nw $(tail -1 $O/g1_nw$nw.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value_warm_l2'])")"
done
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ --launch-skip 6 -c 3 -o $O/prof_edt -f \
   python scripts/bench_edt.py 256 --no-ref) > $O/ncu_edt.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_traj_dyn -s 3 -c 1 -o $O/prof_traj_dyn -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept_dynamics --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --sharded 0 --reference-design 0 --extra-workloads "") > $O/ncu_dyn.log 2>&1
tail -5 $O/tests.log; tail -2 $O/memcheck.log; tail -2 $O/racecheck.log; tail -2 $O/dyn_memcheck.log; tail -2 $O/dyn_racecheck.log; cat $O/edt_bench.jsonl | cut -c1-400
