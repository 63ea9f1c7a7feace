#!/bin/bash
# round 2, session j: mesh obstacles on the GPU (parity, sanitizer, timing) + the full suite
mkdir -p gpurun_out/r2j; O=gpurun_out/r2j
(timeout 900 python -m pytest tests/test_gpu_mesh.py -m gpu -q -p no:cacheprovider) > $O/mesh_tests.log 2>&1; tail -3 $O/mesh_tests.log
(timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_mesh.py -m gpu -q -p no:cacheprovider) > $O/mesh_memcheck.log 2>&1; tail -2 $O/mesh_memcheck.log
(timeout 300 python scripts/bench_mesh.py) > $O/mesh_bench.json 2>&1; cat $O/mesh_bench.json | cut -c1-600
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
