#!/bin/bash
# iteration script: the full GPU suite, front-end timing, a short bench line
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q --timeout 120 2>&1 | tail -6 | cut -c1-300
timeout 120 python scripts/logmel_time.py 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 200 > gpurun_out/k1_bench.json 2> gpurun_out/k1_bench.err; tail -3 gpurun_out/k1_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/k1_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "seq", round(d["sequential"]["ms_per_step"], 4), "host_issue", round(d["host_issue_ms_per_step"], 4), "enc", round(d["roofline"]["ms_per_launch"], 4), "fe", round(d["roofline_frontend"]["ms_per_launch"], 4), "frac_fe", round(d["roofline_frontend"]["frac"], 4), "e2e", round(d["e2e"]["value"]), d["e2e"]["matches_device_path"])
PY
