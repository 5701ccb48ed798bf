#!/bin/bash
# usage: gpu_run_scale.sh N
N=$1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 1000 --warmup 10 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); print({k: d[k] for k in ('n_gpus','value','ms_per_step','scaling')}, d['e2e']['value'], d['clocks'])"
tail -3 gpurun_out/bench_n$N.err
