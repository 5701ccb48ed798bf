#!/bin/bash
# One gpurun call: GPU parity tests, smoke, refinement throughput, ncu of the refinement kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 400 python scripts/refine_times.py --reads 4096 --bases 1000 --near-caps 128,256,384,512,1024 --json gpurun_out/refine_times.json > gpurun_out/refine_times.log 2>&1; echo "refine_times rc=$?"; cat gpurun_out/refine_times.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_refine.csv python scripts/refine_times.py --reads 2048 --bases 600 --cpu-seconds 0.05 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:refine_dp -c 1 -f -o gpurun_out/refine_dp python scripts/refine_times.py --reads 2048 --bases 600 --cpu-seconds 0.05 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out
