#!/usr/bin/env python
"""Opcode histogram of an ncu report's source page, weighted by executed warp instructions and stall samples.
usage: sass_hist.py report.ncu-rep [kernel-regex]"""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line: kernel name; second: header
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = csv.DictReader(io.StringIO("\n".join(lines[start:])))
ops = collections.Counter(); samples = collections.Counter(); tot = 0; tots = 0
for row in rd:
    src = row["Source"].strip()
    if not src: continue
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    try:
        n = int(row["Instructions Executed"]); s = int(row["# Samples"])
    except (ValueError, KeyError):
        continue
    base = op.rstrip(";")
    ops[base] += n; samples[base] += s; tot += n; tots += s
print(f"total warp-inst {tot}  samples {tots}")
for op, n in ops.most_common(40):
    print(f"{op:28s} {n:14d} {100*n/tot:6.2f}%   samples {100*samples[op]/max(tots,1):6.2f}%")
