#!/bin/bash
# Round-end evidence on one GPU: all GPU tests, smoke, bench line (with CPU baseline), reference arm, ncu full captures
# of both hot kernels and the launch list.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
timeout 300 python bench.py 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400
tail -3 gpurun_out/bench.err
for a in uit_xxxs uit_xxs; do timeout 120 python bench.py --arch $a --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_$a.json; done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-300
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:encoder_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_encoder $B > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 3 -c 1 -f -o gpurun_out/prof_logmel $B > gpurun_out/ncu_logmel.log 2>&1; echo "ncu logmel $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list $?"
