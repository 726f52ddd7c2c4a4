#!/usr/bin/env python
"""Summarise one `ncu --set full` capture (.ncu-rep) as a small tracked markdown file under profiles/.

    python tools/ncu_summary_md.py gpurun_out/<name>.ncu-rep profiles/<name>.md "<title>" ["<command that produced it>"]

Reads the report with `ncu -i ... --page raw --csv` and `--page source --csv` (SASS level): headline metrics, stall
reasons per issued instruction, opcode mix, and where the warp-stall samples fall along the SASS (in blocks of 100
instructions, which separates prologue / main loop / flush of the histogram kernels).
"""
import collections
import csv
import io
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
cmd = sys.argv[4] if len(sys.argv) > 4 else ""


def page(name):
    txt = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(txt)))


raw = page("raw")
h, u, v = raw[0], raw[1], raw[2]
get = lambda k: (v[h.index(k)], u[h.index(k)]) if k in h else ("n/a", "")
lines = [f"# {title}", ""]
if cmd:
    lines += [f"Capture: `{cmd}` (`ncu --set full --clock-control none --import-source on`; one launch, cold cache).", ""]
lines += [f"Kernel: `{get('Kernel Name')[0]}`", "", "| metric | value |", "|---|---|"]
for k in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
          "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_global_red.sum",
          "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]:
    val, unit = get(k)
    if val != "n/a":
        lines.append(f"| `{k}` | {val} {unit} |")
lines += ["", "Warp stalls per issued instruction (`smsp__average_warps_issue_stalled_*_per_issue_active.ratio`):", "", "| reason | ratio |", "|---|---|"]
st = [(k.split("stalled_")[1].split("_per_issue")[0], float(v[i])) for i, k in enumerate(h)
      if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k]
for name, val in sorted(st, key=lambda t: -t[1])[:9]:
    lines.append(f"| {name} | {val:.3f} |")

src = page("source")
sh = src[1]
si, ni, ei = sh.index("Source"), sh.index("# Samples"), sh.index("Instructions Executed")
cols = {k: sh.index(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_lg", "stall_wait", "stall_math", "stall_branch_resolving") if k in sh}
data = src[2:]
tot_s = sum(int(r[ni] or 0) for r in data) or 1
tot_e = sum(int(r[ei] or 0) for r in data) or 1
ops = collections.Counter()
for r in data:
    t = r[si].split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")).split(".")[0]
    ops[op] += int(r[ei] or 0)
lines += ["", f"Opcode mix ({tot_e} warp instructions, {len(data)} SASS instructions in the kernel):", "",
          ", ".join(f"{op} {100 * c / tot_e:.1f}%" for op, c in ops.most_common(14)), "",
          f"Where the {tot_s} warp-stall samples fall (blocks of 100 SASS instructions with at least 1 % of the samples):", "",
          "| SASS # | samples | executed | dominant stalls |", "|---|---|---|---|"]
for i in range(0, len(data), 100):
    blk = data[i:i + 100]
    s = sum(int(r[ni] or 0) for r in blk)
    if s < 0.01 * tot_s:
        continue
    e = sum(int(r[ei] or 0) for r in blk)
    dom = sorted(((sum(int(r[c] or 0) for r in blk), k) for k, c in cols.items()), reverse=True)[:2]
    lines.append(f"| {i}-{i + len(blk) - 1} | {100 * s / tot_s:.1f}% | {100 * e / tot_e:.1f}% | " + ", ".join(f"{k[6:]} {100 * n / max(s, 1):.0f}%" for n, k in dom if n) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
