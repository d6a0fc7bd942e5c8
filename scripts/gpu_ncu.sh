#!/bin/bash
# ncu evidence for the two hot kernels (1 GPU; numbers printed under ncu are never bench values)
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-}"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_encoder $B > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 3 -c 1 -f -o gpurun_out/prof_logmel $B > gpurun_out/ncu_logmel.log 2>&1; echo "ncu logmel $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_list.log 2>&1; echo "ncu list $?"
ls -la gpurun_out/*.ncu-rep
