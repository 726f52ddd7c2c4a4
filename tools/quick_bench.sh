#!/bin/bash
# tests + bench + launch list (run under gpurun); usage: bash tools/quick_bench.sh <tag>
TAG=$1
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/${TAG}_bench.json
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('value',d['value'], 'ms/step',d['ms_per_step'], 'kernel_ms',d['roofline']['kernel_ms'], 'frac',d['roofline']['frac'], 'e2e',d['e2e']['value'], d['parity'])"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${TAG}_ncu.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
agg=collections.OrderedDict()
for r in rows[1:]:
    key=(r[ki][:50], r[mi]); v=float(r[vi].replace(',',''))
    agg.setdefault(key,[]).append(v)
for (n,m),l in agg.items():
    if 'fill' in n: continue
    print(f"{n:52s} {m:28s} n={len(l):4d} med={sorted(l)[len(l)//2]:14.1f} max={max(l):14.1f}")
PY
