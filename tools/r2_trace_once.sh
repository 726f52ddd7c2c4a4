#!/bin/bash
mkdir -p gpurun_out
XH_BENCH_TRACE=1 timeout 100 python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --e2e-steps 2 > gpurun_out/trace_once.json 2> gpurun_out/trace_once.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/trace_once.json'))
print('ms_per_step', round(d['ms_per_step'], 4), 'value', d['value'], 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
PY
grep "trace" gpurun_out/trace_once.err | head -2 | cut -c1-300
