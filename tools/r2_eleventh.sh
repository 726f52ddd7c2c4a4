#!/bin/bash
V=$PWD/xhistogram_b200/variants
for rep in 1 2; do
for lib in main nostats; do
  if [ $lib = main ]; then unset XHIST_B200_LIB; else export XHIST_B200_LIB=$V/libxhist_b200_$lib.so; fi
  timeout 200 python tools/r2_ab.py 2.5e8 uniform_counts 6
  timeout 200 python tools/r2_ab.py 2.5e8 uniform 6
done
for lib in main pf2 pf4; do
  if [ $lib = main ]; then unset XHIST_B200_LIB; else export XHIST_B200_LIB=$V/libxhist_b200_$lib.so; fi
  timeout 200 python tools/r2_ab.py 1e9 weighted 8
  timeout 200 python tools/r2_ab.py 1e9 counts 8
  timeout 200 python tools/r2_ab.py 1e9 rows 8
done
done
