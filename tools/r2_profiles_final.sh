#!/bin/bash
# final round-2 profile session: full capture of the headline kernel (with the L2 prefetch) and the launch list of the bench command
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -s 2 -c 1 -f -o gpurun_out/r2_prof_headline_prefetch python tools/r2_one_call.py 1e9 weighted 4 > gpurun_out/r2_ncu_headline_prefetch.log 2>&1
tail -1 gpurun_out/r2_ncu_headline_prefetch.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --samples 2.5e8 > gpurun_out/r2_ncu_launch.log 2>&1
tail -1 gpurun_out/r2_ncu_launch.log | cut -c1-200
ls -la gpurun_out | tail -5
