#!/bin/bash
# bench lines of every BASELINE config on one GPU (+ the reference arm); NGPU>1: the sharded lines under torchrun
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
N=${NGPU:-1}
TAG=${TAG:-r2}
run() {   # name, args...
  local name=$1; shift
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py "$@" > gpurun_out/${TAG}_${name}_n1.json 2> gpurun_out/${TAG}_${name}_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}_n$N.json 2> gpurun_out/${TAG}_${name}_n$N.err
  fi
  echo "== $name rc=$?"; tail -c 600 gpurun_out/${TAG}_${name}_n$N.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${name}_n$N.json").read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print("   value", round(d["value"]), d["unit"], "ms/step", round(d["ms_per_step"], 4), "roofline", r.get("kernel"), round(r.get("frac", 0), 4), "ms", round(r.get("ms_per_launch", 0), 4),
          "fe", round(d.get("roofline_frontend", {}).get("ms_per_launch", d.get("frontend_ms_per_launch", 0)), 4), "e2e", round(d["e2e"]["value"]), "ok", d["e2e"].get("matches_device_path"), "launches", d.get("gpu_launches"))
    for k in ("cpu_baseline", "cpu_baseline_1thread", "gpu_eager_comparator", "parity_max_abs_err_vs_oracle"):
        if k in d: print("   ", k, d[k])
except Exception as e:
    print("   parse failed:", e)
PY
}
for c in ${CONFIGS:-headline c2 c3 c4 c5}; do
  run $c --config $c ${BENCH_ARGS:-}
done
if [ "${REFERENCE:-1}" = "1" ]; then
  run reference --impl reference --steps 3 --warmup 1
fi
