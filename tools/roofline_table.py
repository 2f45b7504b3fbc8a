#!/usr/bin/env python
"""Markdown table of the hardware roofline figures in profiles/summary.json (written by tools/ncu_op_summary.py from the
`ncu --set full` captures of scripts/prof_one.py): per captured operation the dominant kernel, the IMAD.WIDE multiply-
accumulates it executed per row, its duration under ncu, the resulting rate against the measured (peaks_int.json) and the
nominal (32 lanes/clk/SM) multiplier peak, and the step-level figure over every kernel of the operation.
usage: roofline_table.py [summary.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
summ = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "summary.json")))
peaks = json.load(open(os.path.join(ROOT, "peaks_int.json")))
peak = float(peaks.get("imad_wide_chip_gmacs", 8703.2))
nominal = 32 * 148 * 1.965
ORDER = ["verify_k256", "verify_k256_rowpath", "verify_p256", "verify_p256_rowpath", "mul_gen_k256", "mul_var_k256", "mul_var_k256_ct", "mul_var_p384", "mul_var_sm2"]
print("| operation (rows) | dominant kernel | regs | kernel ms | MAC/row (kernel / step) | kernel Tmac/s | of measured / nominal peak | step-level | `fmaheavy` busy | issue active | DRAM GB per call |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for key in ORDER:
    v = summ.get(key)
    if not isinstance(v, dict) or "kernels" not in v:
        continue
    n = v["n_rows"]
    dom = max(v["kernels"], key=lambda k: k["ms"])
    macs_dom = dom["wide_warp_inst"] * 32.0 / n
    rate = n * macs_dom / (dom["ms"] * 1e-3) / 1e9
    step = n * v["wide_macs_per_row"] / (v["gpu_time_ms"] * 1e-3) / 1e9
    print("| `%s` (2^%d) | `%s` | %s | %.2f of %.2f | %.0f / %.0f | %.2f | %.3f / %.3f | %.3f | %.1f %% | %.1f %% | %.1f |" % (
        key, n.bit_length() - 1, dom["kernel"], v.get("registers"), dom["ms"], v["gpu_time_ms"], macs_dom, v["wide_macs_per_row"], rate / 1e3,
        rate / peak, rate / nominal, step / peak, dom.get("fmaheavy_pct") or 0, dom.get("issue_active_pct") or v.get("issue_active_pct") or 0,
        v["dram_bytes_per_launch"] / 1e9))
