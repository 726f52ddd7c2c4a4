#!/bin/bash
# weight lists on device-resident inputs: one fused pass per weight array (default) vs the one-pass kernel (XH_FLAG_ONE_PASS);
# then the driver's three steps on the tree
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "list_of_weights or weighted_mean" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_final_bench_n1.json'))
print(d['value'], d['roofline']['frac'])
for r in d['configs']:
    if 'TWO weight' in r['config']:
        print({k: r[k] for k in ('list_call_ms', 'one_pass_kernel_ms', 'two_calls_ms', 'frac', 'max_rel_diff_vs_separate_calls', 'max_rel_diff_one_pass_kernel')}, r['host_inputs'])
PY
tail -2 gpurun_out/r2_final_bench_n1.err
