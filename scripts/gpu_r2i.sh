#!/bin/bash
# round 2, session i: the driver's commands -- full GPU suite, smoke, bench at N=1 (both arms) and N=2
mkdir -p gpurun_out/r2i; O=gpurun_out/r2i
(timeout 300 python __graft_entry__.py --smoke) > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
(timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider) > $O/gpu_tests.log 2>&1
(timeout 900 python bench.py --steps 20 --warmup 3) > $O/bench_n1.log 2>&1
(timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > $O/bench_ref.log 2>&1
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  (timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3) > $O/bench_n2.log 2>&1
  tail -1 $O/bench_n2.log | cut -c1-2500
fi
tail -2 $O/smoke.log; tail -3 $O/gpu_tests.log; tail -1 $O/bench_n1.log | cut -c1-9000
