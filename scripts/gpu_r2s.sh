#!/bin/bash
# round 2, session s: team kernel without the overflow path (full-length list segments): tests + sweep for the selection rule
mkdir -p gpurun_out/r2s; O=gpurun_out/r2s
timeout 600 python -m pytest tests/test_gpu_rollout.py -x -q -m gpu -k "team_kernel or big_robot" 2>&1 | tail -3 | tee $O/tests.log
timeout 600 python scripts/bench_team.py g1_29_8192_esdf 128,512,1024,1536,2048,3072,4096,8192 2>&1 | tee $O/sweep.log
timeout 600 python scripts/bench_team.py g1_43_8192_esdf 128,512,1024,2048,4096,8192 2>&1 | tee -a $O/sweep.log
timeout 300 python scripts/bench_team.py franka_16384_esdf 256,1024,2048,4096 2>&1 | tee -a $O/sweep.log
