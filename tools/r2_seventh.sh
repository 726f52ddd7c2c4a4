#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in dyn nodyn dyn nodyn; do
  if [ $v = nodyn ]; then export XHIST_B200_LIB=$PWD/xhistogram_b200/variants/libxhist_b200_nodyn.so; else unset XHIST_B200_LIB; fi
  echo "== $v"
  timeout 600 python tools/r2_overheads.py --reps 20 2>&1 | grep -E "cfg3 n=1e\+09|cfg3 n=1.25e\+08" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['case'], 'wall', round(d['wall_ms_med'],4), 'kernel', round(d['kernel_ms_med'],4), round(d['kernel_ms_min'],4), {k:round(v,1) for k,v in d['phases_us_med'].items()})"
done
unset XHIST_B200_LIB
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -s 2 -c 1 -f -o gpurun_out/r2h_prof_shard python tools/r2_one_call.py 1.25e8 weighted 4 > gpurun_out/r2h_ncu_shard.log 2>&1
tail -2 gpurun_out/r2h_ncu_shard.log
