#!/bin/bash
# round 2, session q: team kernel with full-size list segments; phase ablation at small batch
mkdir -p gpurun_out/r2q; O=gpurun_out/r2q
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "team_kernel" 2>&1 | tail -3 | tee $O/tests.log
timeout 600 python scripts/ablate_team.py g1_29_8192_esdf 128,1024 2>&1 | tee $O/ablate.log
timeout 600 python scripts/bench_team.py g1_29_8192_esdf 128,512,1024,2048 2>&1 | tee $O/sweep.log
