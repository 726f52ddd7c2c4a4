#!/bin/bash
# A/B: k_hist compiled for 768 vs 1024 threads per CTA (run under gpurun)
for T in 768 1024; do
  make -C xhistogram_b200/csrc clean >/dev/null; make -C xhistogram_b200/csrc -j4 EXTRA=-DXHK_THREADS=$T 2>&1 | grep -E "error" 
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ab_$T.json
  python -c "
import json; d=json.load(open('gpurun_out/ab_$T.json')); print($T, d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
make -C xhistogram_b200/csrc clean >/dev/null; make -C xhistogram_b200/csrc -j4 2>&1 | grep -E "error"
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r1i_ncu.log 2>&1
