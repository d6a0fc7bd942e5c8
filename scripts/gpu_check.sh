#!/bin/bash
# Run on the GPU box through gpurun: tests, smoke, a short bench and the ncu launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS:-} 2> gpurun_out/bench.err | tee gpurun_out/bench.json
tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --batch 4096 --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/launches.csv
