#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/sanitize_step.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|error" gpurun_out/sanitize_$tool.log | head -20
done
timeout 600 python scripts/pipeline_times.py --reads 256 --bases 2000 --json gpurun_out/pipeline_times.json > gpurun_out/pipeline_times.log 2>&1; echo "pipeline rc=$?"; cat gpurun_out/pipeline_times.log | tail -8
timeout 600 python -m pytest tests/test_io.py -m gpu -x -q 2>&1 | tail -3
