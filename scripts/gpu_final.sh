#!/bin/bash
# Round-end evidence on one GPU: all GPU tests, smoke, bench line (with CPU baseline and comparators), reference arm, the other
# BASELINE configs, ncu full captures of both hot kernels and the launch list.  Everything lands in gpurun_out/ (TAG prefix).
set -u
TAG=${TAG:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -4 | tee gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py 2> gpurun_out/${TAG}_bench.err | tee gpurun_out/${TAG}_bench.json | cut -c1-300
tail -3 gpurun_out/${TAG}_bench.err
for c in c2 c3 c4 c5; do timeout 300 python bench.py --config $c 2>/dev/null > gpurun_out/${TAG}_${c}_n1.json; echo "$c rc=$?"; done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 2> gpurun_out/${TAG}_bench_ref.err | tee gpurun_out/${TAG}_reference_arm.json | cut -c1-300
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-pipeline"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:encoder_tc_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_encoder $B > gpurun_out/ncu_enc.log 2>&1; echo "ncu enc $?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:logmel_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_logmel $B > gpurun_out/ncu_logmel.log 2>&1; echo "ncu logmel $?"
timeout 200 ncu --set full --clock-control none -k regex:head_tc_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_head $B > gpurun_out/ncu_head.log 2>&1; echo "ncu head $?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_list.log 2>&1; echo "ncu list $?"
