#!/bin/bash
# stage timeline only (trace build), then restore the normal build
set -u
mkdir -p gpurun_out
UITK_TRACE=1 python -m uit_mobile_b200.build --force > gpurun_out/build_trace.log 2>&1 || { echo TRACE BUILD FAILED; tail -20 gpurun_out/build_trace.log; exit 1; }
timeout 90 python scripts/tc_trace.py 2>&1 | head -${TRACE_LINES:-60}
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1
