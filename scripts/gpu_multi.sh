#!/bin/bash
# 2-GPU validation: sharded bench under torchrun (NCCL) + reference arm
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
N=${NGPU:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
tail -5 gpurun_out/bench_n$N.err
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/shard_invariance.py 2>&1 | tail -5 | tee gpurun_out/shard_invariance_n$N.log

