#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python scripts/refine_times.py --reads 4096 --bases 1000 --cpu-seconds 2 --json gpurun_out/refine_times.json > gpurun_out/refine_times.log 2>&1; echo "refine rc=$?"; grep "\[gpu\]" gpurun_out/refine_times.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], {k: (v.get('value'), v.get('roofline',{}).get('frac')) for k,v in d['next_rows'].items()})"
