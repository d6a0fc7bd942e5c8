#!/bin/bash
# In-kernel stage timeline of encoder_tc_kernel (trace build), then restore the normal build.
set -u
mkdir -p gpurun_out
UITK_TRACE=1 python -m uit_mobile_b200.build --force > gpurun_out/build_trace.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build_trace.log; exit 1; }
timeout 300 python scripts/tc_trace.py 2>&1 | tail -80
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1
