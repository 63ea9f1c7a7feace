#!/bin/bash
# round 2, session w: the driver's N = 8 / 4 / 2 commands on the final build
mkdir -p gpurun_out/r3j; O=gpurun_out/r3j
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
for n in 8 4 2; do
  if [ "$NG" -ge "$n" ]; then
    (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 3) > $O/bench_n$n.log 2>&1
    tail -1 $O/bench_n$n.log > $O/bench_n${n}_line.json
    python -c "import json,sys; d=json.load(open('$O/bench_n${n}_line.json')); print('N', d['n_gpus'], 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'sharded', {k: (round(v.get('rollout_ms_per_step',-1),4), round(v.get('solve_ms',-1),2)) if 'skipped' not in v else 'skipped' for k,v in d['sharded'].items()})" || tail -5 $O/bench_n$n.log
  fi
done
