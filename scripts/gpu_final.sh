#!/bin/bash
# Final GPU session of the round: everything gpu_round.sh runs without the rollout ncu captures (unchanged since r20),
# plus the RNEA timing against the reference kernels and one ncu --set full capture of the two RNEA CTA kernels.
bash scripts/gpu_round.sh noprof
(timeout 200 python scripts/bench_dynamics.py) > gpurun_out/rnea_bench.jsonl 2>&1
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:rnea_forward_cta --launch-skip 5 -c 1 -o gpurun_out/prof_rnea_fwd -f \
   python scripts/bench_dynamics.py --only franka --no-ref) > gpurun_out/ncu_rnea_fwd.log 2>&1
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:rnea_backward_cta --launch-skip 5 -c 1 -o gpurun_out/prof_rnea_bwd -f \
   python scripts/bench_dynamics.py --only franka --no-ref) > gpurun_out/ncu_rnea_bwd.log 2>&1
cut -c1-330 gpurun_out/rnea_bench.jsonl
