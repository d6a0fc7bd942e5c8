#!/bin/bash
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:encoder_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_encoder python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc $?"
