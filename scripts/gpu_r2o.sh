#!/bin/bash
# round 2, session o: the driver's scaling commands at N = 4 and N = 8 (whatever the box has)
mkdir -p gpurun_out/r2o; O=gpurun_out/r2o
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
for n in 8 4; do
  if [ "$NG" -ge "$n" ]; then
    (timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 3) > $O/bench_n$n.log 2>&1
    tail -1 $O/bench_n$n.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'sharded', {k: (round(v.get('rollout_ms_per_step',-1),4), round(v.get('solve_ms',-1),2)) if 'skipped' not in v else 'skipped' for k,v in d['sharded'].items()})" || tail -5 $O/bench_n$n.log
    (timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --impl reference --gpus $n --steps 5 --warmup 3) > $O/bench_ref_n$n.log 2>&1; tail -1 $O/bench_ref_n$n.log | cut -c1-200
  fi
done
