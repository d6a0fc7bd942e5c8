#!/bin/bash
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 180 python scripts/tc_debug.py 2>&1 | tail -16 | tee gpurun_out/tc_debug.log
echo "tc_debug exit ${PIPESTATUS[0]}"
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_tc.log
timeout 300 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:-} 2> gpurun_out/bench_tc.err | tee gpurun_out/bench_tc.json
tail -3 gpurun_out/bench_tc.err
