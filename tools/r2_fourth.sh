#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 python tools/r2_overheads.py --reps 20 > gpurun_out/r2d_overheads.log 2>&1
grep -E "cfg3 n=|cfg3-counts n=" gpurun_out/r2d_overheads.log
