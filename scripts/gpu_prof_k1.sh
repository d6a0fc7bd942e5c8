#!/bin/bash
# ncu --set full capture of logmel_kernel (one launch, 4096 x 1 s clips) -> gpurun_out/prof_logmel.ncu-rep
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 3 -c 1 -f -o gpurun_out/prof_logmel python scripts/logmel_time.py > gpurun_out/ncu_logmel.log 2>&1; echo "ncu logmel $?"
tail -3 gpurun_out/ncu_logmel.log
