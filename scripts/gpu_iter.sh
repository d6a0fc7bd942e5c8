#!/bin/bash
# One development iteration on the GPU box: tensor-core tests, bench line, then the stage timeline (trace build).
set -u
mkdir -p gpurun_out
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 150 python -m pytest tests/test_gpu_tensorcore.py -x -q --timeout 60 2>&1 | tail -6 | tee gpurun_out/pytest_tc.log
timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${BENCH_ARGS:-} 2> gpurun_out/bench_iter.err | tee gpurun_out/bench_iter.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'enc ms', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],4), 'logmel ms', round(d['roofline_frontend']['ms_per_launch'],4), 'e2e', round(d['e2e']['value']), 'parity', d.get('parity_max_abs_err_vs_oracle'))
"
tail -3 gpurun_out/bench_iter.err
if [ "${TRACE:-1}" = "1" ]; then
UITK_TRACE=1 python -m uit_mobile_b200.build --force > gpurun_out/build_trace.log 2>&1 || { echo TRACE BUILD FAILED; exit 1; }
timeout 90 python scripts/tc_trace.py 2>&1 | tail -60
python -m uit_mobile_b200.build --force > gpurun_out/build.log 2>&1
fi
