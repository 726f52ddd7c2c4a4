#!/bin/bash
# does the nvidia-smi poller perturb the timed steps?  per-step wall times of the headline loop at three polling periods
mkdir -p gpurun_out
for ms in 50 200 50 200; do
  XH_BENCH_SMI_MS=$ms XH_BENCH_TRACE=1 timeout 120 python bench.py --steps 20 --warmup 3 --no-cpu --no-configs --e2e-steps 1 > gpurun_out/smi_$ms.json 2> gpurun_out/smi_$ms.err
  python - <<PY
import json
d = json.load(open('gpurun_out/smi_$ms.json'))
print('period $ms ms: ms_per_step', round(d['ms_per_step'], 4), 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'samples', d['clocks']['samples'])
PY
  grep "trace" gpurun_out/smi_$ms.err | head -1 | cut -c1-400
done
