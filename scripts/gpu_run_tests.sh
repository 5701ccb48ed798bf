#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python scripts/refine_times.py --reads 4096 --bases 1000 --json gpurun_out/refine_times.json > gpurun_out/refine_times.log 2>&1; echo "refine_times rc=$?"; cat gpurun_out/refine_times.log
