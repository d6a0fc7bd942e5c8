#!/bin/bash
# all GPU tests (+ optional short bench): bash scripts/gpu_tests.sh [pytest args]
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -s "$@" 2>&1 | grep -E "max\|d|passed|failed|Error|error|assert|FAILED|^E " | tail -60 | tee gpurun_out/pytest_gpu.log
