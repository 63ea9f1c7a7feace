#!/bin/bash
# round 2, session p: team kernel (TEAM warps per row) -- parity tests, then the small-batch sweep
mkdir -p gpurun_out/r2p; O=gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "team_kernel or big_robot" 2>&1 | tail -5 | tee $O/tests.log
timeout 600 python scripts/bench_team.py g1_29_8192_esdf,g1_43_8192_esdf 128,256,512,1024,1536,2048,4096 2>&1 | tee $O/sweep.log
timeout 300 python scripts/bench_team.py franka_16384_esdf 256,1024,2048 2>&1 | tee -a $O/sweep.log
