#!/bin/bash
# What the driver runs at round end (GPU tests, smoke, both bench arms) + the launch list of the bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['next_rows']['signal_mapping_refinement'].get('value'))"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python scripts/conv_times.py 2>&1 | grep -E "tiled|layers" | tee gpurun_out/conv_times.log
timeout 400 python scripts/refine_times.py --reads 4096 --bases 1000 --cpu-seconds 2 --json gpurun_out/refine_times.json > gpurun_out/refine_times.log 2>&1; grep "\[gpu\]" gpurun_out/refine_times.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
