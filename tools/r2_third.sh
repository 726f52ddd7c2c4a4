#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<double" -s 2 -c 1 -f -o gpurun_out/r2b_prof_cfg5 python tools/cfg5_once.py 1e8 > gpurun_out/r2b_ncu_cfg5.log 2>&1
tail -2 gpurun_out/r2b_ncu_cfg5.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -s 2 -c 1 -f -o gpurun_out/r2b_prof_shard python tools/r2_one_call.py 1.25e8 weighted 4 > gpurun_out/r2b_ncu_shard.log 2>&1
tail -2 gpurun_out/r2b_ncu_shard.log
ls -la gpurun_out/
