#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/r2_overheads.py --reps 20 > gpurun_out/r2f_overheads.log 2>&1
grep -E "cfg3 n=|cfg3-counts n=" gpurun_out/r2f_overheads.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -c 7000 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
