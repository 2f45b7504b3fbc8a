#!/bin/bash
# A/B of library builds on one GPU box: per-kernel times from an ncu launch list of one verify batch per curve, then
# scripts/quick_bench.py.  usage: ab_variants.sh <name> ...   ("main" = the in-tree libecb200.so, else variants/libecb200_<name>.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = main ]; then unset ECB200_LIB; else export ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so; fi
  for c in "k256 22" "p256 20"; do
    set -- $c
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ab_${v}_$1.csv python scripts/prof_one.py $1 verify $2 1 > /dev/null 2>&1
    echo "== $v $1 2^$2 (ncu per-launch us)"
    python - gpurun_out/ab_${v}_$1.csv <<'PY'
import csv, sys
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    if len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    k = d["Kernel Name"].split("(")[0][-40:]
    if any(s in k for s in ("prep", "wintab", "verify_main", "normalize")):
        print("   %-42s %12.1f" % (k, float(d["Metric Value"].replace(",", "")) / 1e3))
PY
  done
  echo "== $v quick_bench"
  timeout 200 python scripts/quick_bench.py k256,p256 22 2>&1 | grep -E "verify|mul_var ct=0" 
done
