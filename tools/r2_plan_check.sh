#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "round2 or dropin or glue" 2>&1 | tail -3
timeout 300 python tools/r2_overheads.py --reps 20 2>&1 | grep -E "cfg3 n=1e\+09|cfg3 n=1.25e\+08|cfg3 n=1e\+06" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], 'wall', round(d['wall_ms_med'],4), 'kernel', round(d['kernel_ms_med'],4), {k:round(v,1) for k,v in d['phases_us_med'].items()})"
