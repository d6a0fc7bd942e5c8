#!/bin/bash
# N-GPU validation (NGPU=2|4|8): shard invariance (NCCL gather, PeerGather, BatchPipeline), sharded bench lines, reference arm
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; exit 1; }
N=${NGPU:-2}
TAG=${TAG:-r2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29512 scripts/shard_invariance.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_shard_invariance_n$N.log
for c in ${CONFIGS:-headline c3 c5}; do
  extra=""; [ "$c" = "headline" ] && extra="--steps 100"
  timeout 400 $TR --master-port 29513 bench.py --gpus $N --config $c $extra --no-extras > gpurun_out/${TAG}_${c}_n$N.json 2> gpurun_out/${TAG}_${c}_n$N.err
  echo "== $c rc=$?"; tail -2 gpurun_out/${TAG}_${c}_n$N.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${c}_n$N.json").read().strip().splitlines()[-1])
    print("   value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "seq", round(d["sequential"]["ms_per_step"], 4), "enc", round(d["roofline"]["ms_per_launch"], 4),
          "fe", round(d.get("roofline_frontend", {}).get("ms_per_launch", d.get("frontend_ms_per_launch", 0)), 4), "e2e", round(d["e2e"]["value"]), d["e2e"].get("matches_device_path"),
          "h2d floor", round(d["e2e"]["h2d_floor_gbs_per_gpu"], 1), "|", d.get("score_gather"))
except Exception as e:
    print("   parse failed:", e)
PY
done
if [ "${NCCL_COMPARE:-0}" = "1" ]; then
  timeout 400 $TR --master-port 29514 bench.py --gpus $N --steps 100 --no-extras --nccl-gather > gpurun_out/${TAG}_headline_nccl_n$N.json 2> gpurun_out/${TAG}_headline_nccl_n$N.err
  python -c "import json; d=json.loads(open('gpurun_out/${TAG}_headline_nccl_n$N.json').read().strip().splitlines()[-1]); print('   nccl-gather headline: value', round(d['value']), 'ms/step', round(d['ms_per_step'],4))"
fi
