#!/bin/bash
# round-2 first contact: GPU tests on the regenerated goldens + short bench
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -s 2>&1 | grep -E "max\|d prob\||passed|failed|Error|error|assert" | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; tail -c 1500 gpurun_out/bench_first.json
