#!/bin/bash
# round 2, N GPUs: NCCL / peer-store multi-GPU tests, bench at N with the exchange modes (driver form: 20 steps)
mkdir -p gpurun_out
N=${1:-2}
MODES=${2:-"p2p p2pi nccl"}
STEPS=${3:-"20 1000"}
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multigpu.log 2>&1; echo "pytest multigpu rc=$?"; tail -5 gpurun_out/pytest_multigpu.log
fi
for mode in $MODES; do
  for steps in $STEPS; do
    extra="--gather p2p"; [ $mode = nccl ] && extra="--gather nccl"; [ $mode = p2pi ] && extra="--gather p2p --immediate"
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps $steps --warmup 5 $extra > gpurun_out/bench_n${N}_${mode}_s${steps}.json 2> gpurun_out/bench_n${N}_${mode}_s${steps}.err; echo "bench N=$N $mode steps=$steps rc=$?"
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_n${N}_${mode}_s${steps}.json') if l.startswith('{')][-1]); print({k: d.get(k) for k in ('n_gpus','value','ms_per_step','gather_verified')}, d['config'].get('gather'), d['config'].get('gather_deferred'), 'e2e', d['e2e']['value'])"
    tail -2 gpurun_out/bench_n${N}_${mode}_s${steps}.err
  done
done
