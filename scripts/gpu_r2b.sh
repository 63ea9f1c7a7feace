#!/bin/bash
# round 2, session b: new optimizer-level tests, the reworked bench line at N=1 (and N=2 when two GPUs are visible)
mkdir -p gpurun_out/r2b; O=gpurun_out/r2b
(timeout 600 python -m pytest tests/test_gpu_sharded_solve.py -m gpu -q -p no:cacheprovider) > $O/tests.log 2>&1
(timeout 900 python bench.py --steps 20 --warmup 3) > $O/bench_20.log 2>&1
(timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > $O/bench_ref.log 2>&1
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  (timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3) > $O/bench_n2.log 2>&1
  tail -1 $O/bench_n2.log | cut -c1-3000
fi
tail -5 $O/tests.log; tail -1 $O/bench_20.log | cut -c1-7000; tail -1 $O/bench_ref.log | cut -c1-1200
