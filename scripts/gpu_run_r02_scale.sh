#!/bin/bash
# N-GPU bench in the driver's form (20 steps) and at 1000 steps, default exchange (one step behind the compute),
# with per-rank diagnostic windows on stderr
N=${1:-2}
mkdir -p gpurun_out
for steps in 20 1000; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $steps --warmup 5 --diag 3 > gpurun_out/bench_n${N}_s${steps}.json 2> gpurun_out/bench_n${N}_s${steps}.err; echo "bench N=$N steps=$steps rc=$?"
  python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}_s${steps}.json') if l.startswith('{')][-1]); print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','gather_verified')}, d['config'].get('gather'), d['config'].get('gather_deferred'), 'e2e', d['e2e']['value'])"
  grep diag gpurun_out/bench_n${N}_s${steps}.err | sort | head -24
done
