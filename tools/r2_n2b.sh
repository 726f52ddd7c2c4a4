#!/bin/bash
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29537 tools/r2_cfg5_multi.py 5e7 > /dev/null 2>&1
XH_NO_P2P=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29538 tools/r2_cfg5_multi.py 5e7 > /dev/null 2>&1
grep -h "reduction\|local" gpurun_out/r2k_p2p_rank*.log | sort | cut -c1-130
grep -h "reduction" gpurun_out/r2k_nccl_rank*.log | sort | cut -c1-130
