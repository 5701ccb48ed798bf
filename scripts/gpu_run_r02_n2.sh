#!/bin/bash
# round 2, two GPUs: NCCL / peer-store multi-GPU tests, bench at N=2 with both exchange modes (driver form: 20 steps)
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"; tail -5 gpurun_out/pytest_multigpu.log
for mode in p2p nccl; do
  for steps in 20 1000; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $steps --warmup 5 --gather $mode > gpurun_out/bench_n${N}_${mode}_s${steps}.json 2> gpurun_out/bench_n${N}_${mode}_s${steps}.err; echo "bench N=$N $mode steps=$steps rc=$?"
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}_${mode}_s${steps}.json') if l.startswith('{')][-1]); print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','gather_verified')}, d['config'].get('gather'), 'e2e', d['e2e']['value'])"
    tail -2 gpurun_out/bench_n${N}_${mode}_s${steps}.err
  done
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1_s20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_n1_s20.json')); print('N=1 s20', d['value'], d['e2e']['value'])"
