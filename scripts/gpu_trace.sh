#!/bin/bash
# stage timeline only (trace build), then restore the normal build.  TRACE_WARM=N: N back-to-back forwards before the traced one
set -u
mkdir -p gpurun_out
UITK_TRACE=${TRACE_LEVEL:-1} python -m uit_mobile_b200.build --force > gpurun_out/build_trace.log 2>&1 || { echo TRACE BUILD FAILED; tail -20 gpurun_out/build_trace.log; exit 1; }
for w in ${TRACE_WARMS:-3}; do
  echo "--- TRACE_WARM=$w"
  TRACE_WARM=$w timeout 120 python scripts/tc_trace.py 2>&1 | head -${TRACE_LINES:-60}
done
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1
