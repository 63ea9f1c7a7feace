#!/bin/bash
# round 2, session e: banded EDT + the reference's call sites over the b200 backend, on the GPU
mkdir -p gpurun_out/r2e; O=gpurun_out/r2e
(timeout 600 python -m pytest tests/test_gpu_zz_edt.py tests/test_gpu_reference_callsites.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1
(timeout 300 compute-sanitizer --tool memcheck python scripts/bench_edt.py 64 --no-ref) > $O/memcheck.log 2>&1
(timeout 300 compute-sanitizer --tool racecheck python scripts/bench_edt.py 64 --no-ref) > $O/racecheck.log 2>&1
(timeout 300 python scripts/bench_edt.py 128 256) > $O/edt_bench.jsonl 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:edt_ --launch-skip 6 -c 3 -o $O/prof_edt -f \
   python scripts/bench_edt.py 256 --no-ref) > $O/ncu_edt.log 2>&1
tail -5 $O/tests.log; tail -2 $O/memcheck.log; tail -2 $O/racecheck.log; cat $O/edt_bench.jsonl | cut -c1-400
