#!/usr/bin/env python
"""Summarise one kernel of an ncu report (--set full) as a markdown table + opcode histogram.
usage: ncu_summary.py report.ncu-rep [out.md] [title] [kernel-name regex, for reports holding several kernels]"""
import csv, io, subprocess, sys, json, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else None
title = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
kfilter = ["--kernel-name", "regex:" + sys.argv[4]] if len(sys.argv) > 4 else []
raw = subprocess.run(["ncu", "-i", rep] + kfilter + ["--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
WANT = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "sm__cycles_elapsed.avg",
    "launch__local_memory_size" , "smsp__inst_executed_op_local_ld.sum",
]
lines = ["# ncu --set full: %s" % title, "", "| metric | unit | value |", "|---|---|---|"]
for k in WANT:
    if k in m:
        lines.append("| %s | %s | %s |" % (k, m[k][0], m[k][1]))
# opcode histogram
src = subprocess.run(["ncu", "-i", rep] + kfilter + ["--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(src) if l.startswith('"Address"'))
import collections
ops = collections.Counter(); tot = 0
# a report captured with --import-source on lists the kernel twice (SASS view, then source-correlated view): first table only
end = next((i for i in range(start + 1, len(src)) if src[i].startswith('"Address"')), len(src))
for row in csv.DictReader(io.StringIO("\n".join(src[start:end]))):
    s = row["Source"].strip()
    if not s: continue
    t = s.split(); op = t[1] if t[0].startswith("@") else t[0]
    try: n = int(row["Instructions Executed"])
    except (ValueError, TypeError): continue
    ops[op.rstrip(";")] += n; tot += n
lines += ["", "Executed warp instructions by opcode (top 16 of %d):" % tot, "", "| opcode | warp-inst | share |", "|---|---|---|"]
for op, n in ops.most_common(16):
    lines.append("| %s | %d | %.2f %% |" % (op, n, 100.0 * n / tot))
wide = sum(n for op, n in ops.items() if op.startswith("IMAD.WIDE"))
lines += ["", "IMAD.WIDE warp-instructions: %d (%.2f %% of issued); x32 lanes = %.4g multiply-accumulates per launch" % (wide, 100.0 * wide / tot, wide * 32.0)]
text = "\n".join(lines) + "\n"
if out:
    open(out, "w").write(text)
print(text)
