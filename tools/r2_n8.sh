#!/bin/bash
# 8-GPU session: the driver's scaling command at N=8 (and N=4)
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29527 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2j_bench_n$N.json 2> gpurun_out/r2j_bench_n$N.err
tail -c 1500 gpurun_out/r2j_bench_n$N.json; grep -v "^W1017\|^\[W\|^\*\*\*\|OMP_NUM" gpurun_out/r2j_bench_n$N.err | tail -8
