#!/bin/bash
# final single-GPU validation: the driver's three steps (GPU tests, smoke, bench) + the reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
tail -c 200 gpurun_out/r2_final_bench_n1.json; tail -2 gpurun_out/r2_final_bench_n1.err
