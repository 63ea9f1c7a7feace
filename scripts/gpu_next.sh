#!/bin/bash
# First GPU session of the next round: the two rows whose GPU tests were written after round 1's GPU budget was spent
# (dynamics-aware STATE cost composition, exact nearest-site transform), their timings, and one ncu capture per new kernel.
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_zx_dynamics_trees.py tests/test_gpu_zy_effort_cost.py tests/test_gpu_zz_edt.py -m gpu -q -p no:cacheprovider) > gpurun_out/new_rows_tests.log 2>&1
(timeout 300 python scripts/bench_edt.py 128 256) > gpurun_out/edt_bench.jsonl 2>&1
(timeout 300 compute-sanitizer --tool memcheck python scripts/bench_edt.py 64 --no-ref) > gpurun_out/edt_memcheck.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ --launch-skip 6 -c 3 -o gpurun_out/prof_edt -f \
   python scripts/bench_edt.py 256 --no-ref) > gpurun_out/ncu_edt.log 2>&1
tail -5 gpurun_out/new_rows_tests.log; cat gpurun_out/edt_bench.jsonl | cut -c1-300; tail -3 gpurun_out/edt_memcheck.log
