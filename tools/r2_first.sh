#!/bin/bash
# round 2, first GPU session: fixed per-call costs, worst-case data rows, cluster/DSMEM mechanism test
mkdir -p gpurun_out tools/bin
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt
timeout 600 python tools/r2_overheads.py --reps 20 > gpurun_out/r2a_overheads.log 2>&1
cat gpurun_out/r2a_overheads.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/microbench_cluster tools/microbench_cluster.cu 2>&1 | tail -3
timeout 300 tools/bin/microbench_cluster > gpurun_out/r2a_microbench_cluster.log 2>&1
cat gpurun_out/r2a_microbench_cluster.log
timeout 600 python tools/bench_configs.py --reps 5 > gpurun_out/r2a_configs.log 2>&1
tail -15 gpurun_out/r2a_configs.log
