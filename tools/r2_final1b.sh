#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
tail -c 200 gpurun_out/r2v_bench.json; tail -3 gpurun_out/r2v_bench.err
