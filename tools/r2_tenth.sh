#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "golden or random or round2" 2>&1 | tail -3
for v in pipe nopipe pipe nopipe; do
  if [ $v = nopipe ]; then export XHIST_B200_LIB=$PWD/xhistogram_b200/variants/libxhist_b200_nopipe.so; else unset XHIST_B200_LIB; fi
  echo "== $v"; timeout 200 python tools/cfg5_once.py 4e8 | tail -1; timeout 200 python tools/cfg5_once.py 5e7 | tail -1
done
unset XHIST_B200_LIB
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
tail -c 600 gpurun_out/r2n_bench.json; tail -3 gpurun_out/r2n_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-configs --e2e-steps 1 --samples 2.5e8 > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-300
