#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
tail -c 300 gpurun_out/r2u_bench.json; tail -3 gpurun_out/r2u_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2u_bench_ref.json 2> gpurun_out/r2u_bench_ref.err
tail -c 400 gpurun_out/r2u_bench_ref.json
timeout 600 python tools/r2_overheads.py --reps 20 > gpurun_out/r2u_overheads.log 2>&1
