#!/bin/bash
# First GPU session of the next round.  Everything below already passes on the CPU against the kernels compiled for the host SIMT
# emulation (tests/test_emulated_gpu_suite_cpu.py); this is their first contact with a B200:
#   1. the GPU tests written after round 1's last GPU session (RNEA on synthetic trees / every rows-per-CTA variant, the
#      dynamics-aware STATE cost host-composed and inside the trajectory kernel incl. the knots path, the nearest-site transform),
#   2. compute-sanitizer over the new kernels (address-space / out-of-bounds errors are what the emulation cannot see),
#   3. timings: EDT vs the reference's PBA+ kernels, RNEA, the dynamics-aware MPC workloads (bench.py extras),
#   4. one ncu --set full capture per new kernel family.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_zx_dynamics_trees.py tests/test_gpu_zy_effort_cost.py tests/test_gpu_zz_edt.py tests/test_gpu_zzz_center_of_mass.py -m gpu -q -p no:cacheprovider) > gpurun_out/new_rows_tests.log 2>&1
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_zy_effort_cost.py -m gpu -q -p no:cacheprovider -k "dynamics_aware") > gpurun_out/dyn_memcheck.log 2>&1
(timeout 300 compute-sanitizer --tool memcheck python scripts/bench_edt.py 64 --no-ref) > gpurun_out/edt_memcheck.log 2>&1
(timeout 300 python scripts/bench_edt.py 128 256) > gpurun_out/edt_bench.jsonl 2>&1
(timeout 200 python scripts/bench_dynamics.py) > gpurun_out/rnea_bench.jsonl 2>&1
(timeout 900 python bench.py) > gpurun_out/bench.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ --launch-skip 6 -c 3 -o gpurun_out/prof_edt -f \
   python scripts/bench_edt.py 256 --no-ref) > gpurun_out/ncu_edt.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_traj_kernel -s 3 -c 1 -o gpurun_out/prof_traj_dyn -f \
   python bench.py --workload franka_mpc_1024x30_esdf_swept_dynamics --steps 3 --warmup 3 --no-cpu-baseline --ik-solve 0 --rnea 0 --edt 0 --extra-workloads "") > gpurun_out/ncu_traj_dyn.log 2>&1
tail -5 gpurun_out/new_rows_tests.log; tail -3 gpurun_out/dyn_memcheck.log; tail -3 gpurun_out/edt_memcheck.log
cut -c1-300 gpurun_out/edt_bench.jsonl; tail -1 gpurun_out/bench.log | cut -c1-4000
