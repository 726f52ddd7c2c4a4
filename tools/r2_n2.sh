#!/bin/bash
# 2-GPU session: peer-memory reduction, strong-scaling record, multi-rank parity
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err
tail -c 5000 gpurun_out/r2g_bench_n2.json; tail -8 gpurun_out/r2g_bench_n2.err
XH_NO_P2P=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu --no-configs > gpurun_out/r2g_bench_n2_nccl.json 2> gpurun_out/r2g_bench_n2_nccl.err
tail -c 2500 gpurun_out/r2g_bench_n2_nccl.json; tail -3 gpurun_out/r2g_bench_n2_nccl.err
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu or allreduce or dropin" 2>&1 | tail -5
