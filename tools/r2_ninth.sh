#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for s in 4 3 2 1; do echo "== XH_LUT_STEPS=$s"; XH_LUT_STEPS=$s timeout 200 python tools/cfg5_once.py 1e8 | tail -1; XH_LUT_STEPS=$s timeout 200 python tools/cfg5_once.py 4e8 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<double" -s 2 -c 1 -f -o gpurun_out/r2_prof_cfg5_after python tools/cfg5_once.py 1e8 > gpurun_out/r2_ncu_cfg5_after.log 2>&1
tail -1 gpurun_out/r2_ncu_cfg5_after.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -s 2 -c 1 -f -o gpurun_out/r2_prof_headline python tools/r2_one_call.py 1e9 weighted 4 > gpurun_out/r2_ncu_headline.log 2>&1
tail -1 gpurun_out/r2_ncu_headline.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --samples 2.5e8 > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-300
ls -la gpurun_out | tail -8
