#!/bin/bash
# compute-sanitizer over tools/sanitize_case.py (run under gpurun); logs in gpurun_out/
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  echo "== compute-sanitizer --tool $tool python tools/sanitize_case.py"
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  grep -E "err |ok|SUMMARY|hazard|Error|error" gpurun_out/sanitizer_$tool.log | tail -40
done
