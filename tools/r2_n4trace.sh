#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
XH_BENCH_TRACE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2x_bench_n$N.json 2> gpurun_out/r2x_bench_n$N.err
grep "trace rank" gpurun_out/r2x_bench_n$N.err | grep "step5\|step_strong" | cut -c1-400
tail -c 600 gpurun_out/r2x_bench_n$N.json
