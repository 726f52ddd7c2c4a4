#!/bin/bash
# round 2, second GPU session: per-kernel durations of a 1/8-shard call, full captures of the cfg5 kernel and of the shard-size headline kernel
mkdir -p gpurun_out
for n in 1e6 1.25e8; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_launches_$n.csv python tools/r2_one_call.py $n weighted 4 > gpurun_out/r2b_l_$n.log 2>&1
done
python - <<'PY'
import csv
for n in ("1e6", "1.25e8"):
    rows = [r for r in csv.reader(open(f"gpurun_out/r2b_launches_{n}.csv")) if len(r) > 10]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    print("n =", n)
    for r in rows[1:]:
        if "fill" in r[ki]: continue
        print(f"  {r[ki][:70]:72s} {float(r[vi].replace(',',''))/1e3:10.2f} us")
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<double" -s 2 -c 1 -f -o gpurun_out/r2b_prof_cfg5 python tools/cfg5_once.py 1e8 > gpurun_out/r2b_ncu_cfg5.log 2>&1
tail -3 gpurun_out/r2b_ncu_cfg5.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -s 2 -c 1 -f -o gpurun_out/r2b_prof_shard python tools/r2_one_call.py 1.25e8 weighted 4 > gpurun_out/r2b_ncu_shard.log 2>&1
tail -3 gpurun_out/r2b_ncu_shard.log
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -s 2 -c 1 -f -o gpurun_out/r2b_prof_uniform python tools/r2_one_call.py 2.5e8 uniform 4 > gpurun_out/r2b_ncu_uniform.log 2>&1
tail -3 gpurun_out/r2b_ncu_uniform.log
ls -la gpurun_out/r2b*
