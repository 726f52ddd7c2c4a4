#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/r2_overheads.py --reps 20 2>&1 | grep -E "cfg3 n=1e\+09|cfg3 n=1.25e\+08|cfg3-counts n=1" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], 'wall', round(d['wall_ms_med'],4), 'kernel', round(d['kernel_ms_med'],4), round(d['kernel_ms_min'],4), 'frac', round(d['frac_kernel'],3))"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
tail -c 300 gpurun_out/r2r_bench.json; tail -3 gpurun_out/r2r_bench.err
