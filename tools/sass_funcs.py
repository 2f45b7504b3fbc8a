#!/usr/bin/env python
"""Static opcode counts of the device sub-functions (labels reached by CALL) inside one kernel of a cuobjdump -sass listing.
usage: sass_funcs.py file.sass kernel-substring"""
import re, sys, collections
lines = open(sys.argv[1]).read().splitlines()
key = sys.argv[2]
start = next(i for i, l in enumerate(lines) if "Function :" in l and key in l)
end = next((i for i in range(start + 1, len(lines)) if "Function :" in lines[i]), len(lines))
body = lines[start:end]
ins = []
for l in body:
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2)))
# call targets
targets = sorted({int(m.group(1), 16) for a, t in ins for m in [re.search(r"CALL\.\S+\s+.*?0x([0-9a-f]+)", t)] if m})
print("instructions:", len(ins), "call targets:", [hex(t) for t in targets])
bounds = targets + [ins[-1][0] + 16]
def hist(lo, hi):
    c = collections.Counter()
    for a, t in ins:
        if lo <= a < hi:
            toks = t.split()
            op = toks[1] if toks[0].startswith("@") else toks[0]
            c[op] += 1
    return c
if targets:
    print("main body:", sum(hist(0, targets[0]).values()))
for i, t in enumerate(targets):
    h = hist(t, bounds[i + 1])
    # stop at first RET
    print(hex(t), "total", sum(h.values()), dict(h.most_common(14)))
