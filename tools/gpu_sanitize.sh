#!/bin/bash
set -u
mkdir -p gpurun_out
export MPC_RIC_ALWAYS=1
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|riccati |inverse |ticks" gpurun_out/sanitize_$tool.log | head -24
done
