#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/r2_teardown.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
echo "exit ${PIPESTATUS[0]}"
