#!/bin/bash
# Next-round starting point for the erratic one-limb weighted kernel.
# Build the variant first:   touch xhistogram_b200/csrc/xhist_k_f32.cu && make -C xhistogram_b200/csrc EXTRA=-DXH_CHEAP_SIDE_W3=1
# then under gpurun:         bash tools/erratic_probe.sh
# It prints the kernel time of 30 identical config-3 calls and captures 6 consecutive k_hist<float,3,...> launches with
# the metrics that should tell a slow launch from a fast one.  Rebuild WITHOUT the flag afterwards.
mkdir -p gpurun_out
cat > /tmp/erratic.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from xhistogram_b200 import DeviceArray, core
n = int(1e9)
x = DeviceArray.normal((n,), np.float32, seed=3); y = DeviceArray.normal((n,), np.float32, seed=4)
w = DeviceArray.uniform((n,), np.float32, seed=5)
e = np.linspace(-4, 4, 257)
t = {}
for i in range(int(sys.argv[1])):
    core._bincount(x, y, w, weights=True, axis=None, bins=[e, e], _timing=t)
    print(round(t["kernel_ms"], 3), end=" ", flush=True)
print()
PY
python /tmp/erratic.py 30
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
M=$M,smsp__inst_executed_op_shared_atom.sum,smsp__inst_executed_op_global_red.sum,dram__bytes_read.sum
M=$M,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
M=$M,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
timeout 600 ncu --metrics $M --clock-control none --kernel-name-base demangled -k regex:"k_hist<float, .int.3" -c 6 --csv \
    --log-file gpurun_out/erratic_metrics.csv python /tmp/erratic.py 6 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/erratic_metrics.csv')) if len(r) > 10]
h = rows[0]; ii, mi, vi = h.index('ID'), h.index('Metric Name'), h.index('Metric Value')
t = collections.OrderedDict()
for r in rows[1:]:
    t.setdefault(r[mi], {})[r[ii]] = r[vi]
for m, d in t.items():
    print(f"{m[:90]:92s}", "  ".join(f"{v:>14s}" for v in d.values()))
PY
