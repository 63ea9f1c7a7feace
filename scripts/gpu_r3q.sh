#!/bin/bash
# round 2, session 3q: depth -> ESDF chain test after the IEEE-division fix
mkdir -p gpurun_out/r3q; O=gpurun_out/r3q
(timeout 900 python -m pytest tests/test_gpu_zz_edt.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1; tail -25 $O/tests.log | cut -c1-220
