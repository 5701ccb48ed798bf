#!/bin/bash
# round 2, first full pass: GPU tests, smoke, bench (driver form), phase stamps, ncu launch list + full capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
RB200_MEGA_STAMPS=1 timeout 120 python scripts/profile_step.py 6 1024 2> gpurun_out/mega_stamps.log; tail -3 gpurun_out/mega_stamps.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s20.json 2> gpurun_out/bench_s20.err; echo "bench20 rc=$?"; tail -c 600 gpurun_out/bench_s20.err
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], 'cpu', d['cpu_baseline']['value']); print('stock', d.get('gpu_stock_baseline')); print('configs', d.get('configs')); print('variants', d.get('variants'))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mega_kernel -s 2 -c 2 -o gpurun_out/prof_mega -f python scripts/profile_step.py 5 1024 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
