#!/bin/bash
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
for tool in memcheck racecheck synccheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; tail -4 gpurun_out/sanitizer_$tool.log
done
