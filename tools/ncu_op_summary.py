#!/usr/bin/env python
"""Summarise EVERY kernel launch of one operation captured in an ncu report (--set full [--import-source on]) and fold the
totals into profiles/summary.json, the file bench.py reads its hardware roofline numbers from.

usage: ncu_op_summary.py report.ncu-rep out.md key n_rows "title" [--skip-kernel regex]
  key     entry of profiles/summary.json (e.g. mul_var_k256, verify_k256, mul_gen_k256)
  n_rows  rows the captured invocation processed

Per launch: duration, registers, pipe utilisation, top stalls, DRAM bytes, local loads / stores and the executed opcode
histogram (source page).  Totals: IMAD.WIDE multiply-accumulates per row over all launches (x32 lanes), the non-multiply
instructions executed on the FMA pipe, DRAM bytes per invocation, and which launch dominates."""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "launch__local_memory_size" if False else "sm__cycles_elapsed.avg",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3}


def num(v):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    return None if x != x else x       # ncu occasionally reports -nan for a tiny launch: treated as missing


def main():
    rep, out, key, n_rows, title = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5]
    skip = re.compile(sys.argv[sys.argv.index("--skip-kernel") + 1]) if "--skip-kernel" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, launches = rows[0], rows[1], rows[2:]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
    starts = [i for i, l in enumerate(src) if l.startswith('"Address"')]
    tables = []
    for j, st in enumerate(starts):
        end = (starts[j + 1] - 1) if j + 1 < len(starts) else len(src)
        tables.append(src[st:end])
    per = len(tables) // max(len(launches), 1)      # 2 with --import-source on (SASS view, then source view), else 1
    lines = ["# ncu --set full: %s" % title, "", "report: `%s`, %d launches, %d rows" % (os.path.basename(rep), len(launches), n_rows), ""]
    tot_wide = tot_other = tot_inst = 0
    tot_dram = 0.0
    tot_ms = 0.0
    kernels = []
    for li, vals in enumerate(launches):
        m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        name = m["Kernel Name"][1]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("ecb::", "")
        if skip and skip.search(short):
            continue
        ops = collections.Counter()
        if per:
            for row in csv.DictReader(io.StringIO("\n".join(tables[li * per]))):
                s = (row.get("Source") or "").strip()
                if not s:
                    continue
                t = s.split()
                op = (t[1] if t[0].startswith("@") else t[0]).rstrip(";")
                try:
                    ops[op] += int(row["Instructions Executed"])
                except (ValueError, TypeError, KeyError):
                    pass
        ninst = sum(ops.values())
        wide = sum(n for op, n in ops.items() if op.startswith("IMAD.WIDE"))
        other = sum(n for op, n in ops.items() if op.startswith("IMAD") and not op.startswith("IMAD.WIDE"))
        u, v = m["gpu__time_duration.sum"]
        ms = num(v) * UNIT.get(u, 1.0)
        dram = sum(num(m[k][1]) * UNIT.get(m[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m and num(m[k][1]) is not None)
        tot_wide += wide; tot_other += other; tot_inst += ninst; tot_dram += dram; tot_ms += ms
        k = {"kernel": short, "ms": round(ms, 4), "registers": int(num(m["launch__registers_per_thread"][1]) or 0),
             "fmaheavy_pct": num(m["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"][1]),
             "alu_pct": num(m["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"][1]),
             "issue_active_pct": num(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
             "dram_bytes": dram, "wide_warp_inst": wide, "other_fma_pipe_warp_inst": other, "warp_inst": ninst,
             "local_ld": num(m.get("smsp__sass_inst_executed_op_local_ld.sum", ("", "0"))[1]), "local_st": num(m.get("smsp__sass_inst_executed_op_local_st.sum", ("", "0"))[1])}
        kernels.append(k)
        lines += ["## launch %d: `%s`" % (li, short), "", "| metric | unit | value |", "|---|---|---|"]
        for w in WANT:
            if w in m:
                lines.append("| %s | %s | %s |" % (w, m[w][0], m[w][1]))
        if ninst:
            lines += ["", "Executed warp instructions by opcode (top 14 of %d):" % ninst, "", "| opcode | warp-inst | share |", "|---|---|---|"]
            for op, n in ops.most_common(14):
                lines.append("| %s | %d | %.2f %% |" % (op, n, 100.0 * n / ninst))
            lines += ["", "IMAD.WIDE: %d warp-inst (%.2f %% of issued) = %.4g multiply-accumulates; other FMA-pipe integer instructions "
                      "(IMAD.MOV / IMAD.X / IMAD / IMAD.HI ...): %d (%.2f %%)" % (wide, 100.0 * wide / ninst, wide * 32.0, other, 100.0 * other / ninst)]
        lines.append("")
    dom = max(kernels, key=lambda k: k["ms"])
    entry = {"n_rows": n_rows, "kernel": dom["kernel"], "gpu_time_ms": round(tot_ms, 4), "dominant_kernel_ms": dom["ms"],
             "pipe_fmaheavy_pct": dom["fmaheavy_pct"], "pipe_alu_pct": dom["alu_pct"], "issue_active_pct": dom["issue_active_pct"],
             "registers": dom["registers"], "dram_bytes_per_launch": tot_dram, "wide_macs_per_row": round(tot_wide * 32.0 / n_rows, 1),
             "other_fma_pipe_inst_per_row": round(tot_other * 32.0 / n_rows, 1), "warp_inst_total": tot_inst,
             "source": os.path.relpath(out, ROOT), "kernels": kernels}
    lines += ["## totals over the %d launches" % len(kernels), "",
              "* device time %.4f ms (cold-cache, serialised ncu replays; the dominant launch is `%s`, %.4f ms)" % (tot_ms, dom["kernel"], dom["ms"]),
              "* IMAD.WIDE multiply-accumulates per row: %.1f (x32 lanes per warp instruction / %d rows)" % (entry["wide_macs_per_row"], n_rows),
              "* other integer instructions on the FMA pipe per row: %.1f thread-instructions" % entry["other_fma_pipe_inst_per_row"],
              "* DRAM bytes per invocation: %.4g (%.1f B per row)" % (tot_dram, tot_dram / n_rows), ""]
    open(out, "w").write("\n".join(lines))
    sj = os.environ.get("ECB200_SUMMARY_JSON") or os.path.join(ROOT, "profiles", "summary.json")
    try:
        summ = json.load(open(sj))
    except Exception:
        summ = {}
    summ[key] = entry
    json.dump(summ, open(sj, "w"), indent=1)
    print("\n".join(lines[-8:]))


if __name__ == "__main__":
    main()
