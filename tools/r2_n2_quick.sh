#!/bin/bash
# quick 2-GPU sanity of the final tree: peer-memory reduction, strong record, bin-by-bin multi-rank parity (no CPU leg, no config rows)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-configs > gpurun_out/r2_n2_quick.json 2> gpurun_out/r2_n2_quick.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_n2_quick.json'))
print(d['value'], d['ms_per_step'], d.get('parity'), {k: d['strong'][k] for k in d.get('strong', {}) if 'speedup' in k or 'ms' in k}, d['e2e']['value'])
PY
tail -2 gpurun_out/r2_n2_quick.err
