"""Compares the ncu counter files written by scripts/ct_audit.sh: for every (curve, op) the counters of every kernel launch must
be identical for the secret patterns small / random / high / sparse.  Prints a markdown report; exit code 1 on a difference.
usage: ct_audit_compare.py <dir with ct_*.csv>"""
import collections
import csv
import glob
import os
import re
import sys

d = sys.argv[1]
PATTERNS = ["small", "random", "high", "sparse", "small_again"]      # the order scripts/ct_dyn_run.py runs them in
groups = collections.defaultdict(dict)
for f in sorted(glob.glob(os.path.join(d, "ct_*_*.csv"))):
    m = re.match(r"ct_(\w+?)_(mul_var|mul_gen|sign)\.csv", os.path.basename(f))
    if not m:
        continue
    rows = list(csv.reader(l for l in open(f) if l.startswith('"')))
    if not rows:
        continue
    hdr = rows[0]
    per = collections.OrderedDict()
    for r in rows[1:]:
        rec = dict(zip(hdr, r))
        per.setdefault(int(rec["ID"]), {"kernel": re.sub(r"\(.*", "", rec["Kernel Name"]).replace("void ", "")})[rec["Metric Name"]] = rec["Metric Value"]
    ids = sorted(per)
    if len(ids) % len(PATTERNS):
        print("!! %s: %d launches is not a multiple of %d patterns" % (f, len(ids), len(PATTERNS)))
        continue
    q = len(ids) // len(PATTERNS)
    for pi, pat in enumerate(PATTERNS):
        part = collections.OrderedDict()
        for j, i in enumerate(ids[pi * q:(pi + 1) * q]):
            met = dict(per[i])
            part[(j, met.pop("kernel"))] = met
        groups[(m.group(1), m.group(2))][pat] = part
print("# Dynamic constant-time audit: ncu counters of the secret-scalar kernels under four secret patterns\n")
print("Same public inputs (points, prehashes), 2^14 rows, secrets = small (1, 2, 3 ...), random, high (n-1, n-2 ...), sparse (single bits).")
print("`small` is run twice (first and last): a counter that differs between those two IDENTICAL runs is measurement noise of that counter")
print("(listed as noisy, excluded from the verdict).  `identical` means every other counter has the same value under all patterns.\n")
bad = 0
for (curve, op), pats in sorted(groups.items()):
    names = [p for p in PATTERNS if p in pats]
    ref = pats[names[0]]
    print("## %s %s  (patterns: %s)\n" % (curve, op, ", ".join(names)))
    print("| launch | kernel | inst executed | thread inst | divergent branch targets | local ld / st sectors | global ld sectors | shared wavefronts | verdict |")
    print("|---|---|---|---|---|---|---|---|---|")
    for key, met in ref.items():
        diffs = []
        again = pats.get("small_again", {}).get(key, {})
        noisy = sorted(k for k, v in met.items() if again and again.get(k) != v)
        for p in names[1:]:
            if p == "small_again":
                continue
            other = pats[p].get(key)
            if other is None:
                diffs.append("%s: launch missing" % p)
                continue
            for k, v in met.items():
                if other.get(k) != v and k not in noisy:
                    diffs.append("%s: %s %s vs %s" % (p, k, other.get(k), v))
        g = lambda k: met.get(k, "-")
        verdict = ("identical" if not diffs else "DIFFERENT: " + "; ".join(diffs[:4])) + (" (noisy between identical runs: %s)" % ", ".join(n.split("__")[-1] for n in noisy) if noisy else "")
        bad += bool(diffs)
        print("| %s | `%s` | %s | %s | %s | %s / %s | %s | %s | %s |" % (
            key[0], key[1], g("smsp__inst_executed.sum"), g("smsp__thread_inst_executed.sum"), g("smsp__sass_branch_targets_threads_divergent.sum"),
            g("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"), g("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"),
            g("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"), g("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), verdict))
    print()
print("**launches whose counters depend on the secrets: %d**" % bad)
sys.exit(1 if bad else 0)
