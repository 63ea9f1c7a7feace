#!/bin/bash
# round 2, session n: the reference's optimizer over the b200 rollout on the GPU + full suite
mkdir -p gpurun_out/r2n; O=gpurun_out/r2n
(timeout 900 python -m pytest tests/test_gpu_reference_callsites.py -m gpu -q -p no:cacheprovider) > $O/callsites.log 2>&1; tail -5 $O/callsites.log
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/gpu_tests.log 2>&1; tail -3 $O/gpu_tests.log
