#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_io.py -m gpu -x -q > gpurun_out/pytest_io.log 2>&1; echo "pytest io rc=$?"; tail -12 gpurun_out/pytest_io.log
timeout 300 python scripts/vbz_times.py --json gpurun_out/vbz_times.json > gpurun_out/vbz_times.log 2>&1; echo "vbz rc=$?"; tail -5 gpurun_out/vbz_times.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:svb16 -c 1 -f -o gpurun_out/svb16 python scripts/vbz_times.py --reads 256 > gpurun_out/ncu_vbz.log 2>&1; echo "ncu rc=$?"
timeout 300 python scripts/pipeline_times.py --reads 256 --bases 2000 --json gpurun_out/pipeline_times.json > gpurun_out/pipeline_times.log 2>&1; echo "pipeline rc=$?"; tail -5 gpurun_out/pipeline_times.log
