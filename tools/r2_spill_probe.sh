#!/bin/bash
mkdir -p gpurun_out
XH_NO_PACKED=1 timeout 200 python tools/r2_ab.py 2.5e8 uniform_counts 4
XH_NO_PACKED=1 timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.0" -s 2 -c 1 -f -o gpurun_out/r2_prof_spill python tools/r2_ab.py 2.5e8 uniform_counts 3 > gpurun_out/r2_ncu_spill.log 2>&1
tail -1 gpurun_out/r2_ncu_spill.log
