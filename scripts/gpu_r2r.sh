#!/bin/bash
# round 2, session r: team kernel sweep for the selection rule
mkdir -p gpurun_out/r2r; O=gpurun_out/r2r
timeout 600 python scripts/bench_team.py g1_29_8192_esdf 1536,2048,2368,3072,4096,8192 2>&1 | tee $O/sweep.log
timeout 600 python scripts/bench_team.py g1_43_8192_esdf 128,512,1024,2048,4096,8192 2>&1 | tee -a $O/sweep.log
timeout 300 python scripts/bench_team.py franka_16384_esdf 256,1024,2048,4096,16384 2>&1 | tee -a $O/sweep.log
