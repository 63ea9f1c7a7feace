#!/bin/bash
# round 2, session 3o: last check of the committed build: GPU suite + smoke + a short bench line
mkdir -p gpurun_out/r3o; O=gpurun_out/r3o
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -2 $O/gpu_tests.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')") > $O/smoke.log 2>&1; tail -1 $O/smoke.log
(timeout 600 python bench.py --steps 100 --warmup 5 --ik-solve 0 --edt 0 --rnea 0 --sharded 0 --reference-design 0 --no-cpu-baseline --extra-workloads "") > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-260
