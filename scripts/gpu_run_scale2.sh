#!/bin/bash
mkdir -p gpurun_out
run() {  # N G tag
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $1 --steps 1000 --warmup 10 --gather-every $2 > gpurun_out/bench_$3.json 2> gpurun_out/bench_$3.err
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_$3.json') if l.startswith('{')][-1]); print('$3', {k: d[k] for k in ('n_gpus','value','ms_per_step')}, 'e2e', d['e2e']['value'])"
}
run 2 8 n2_g8_a
run 2 1 n2_g1
run 2 8 n2_g8_b
