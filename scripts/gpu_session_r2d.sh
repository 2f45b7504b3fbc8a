#!/bin/bash
# Round-2 GPU session D: per-key tables - window width (4 / 5 / 6 bits) and resident-CTA variants, after the 3-round fill kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for v in main w5 w6 w5c6 w5c5; do
  if [ "$v" = main ]; then unset ECB200_LIB; else export ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so; fi
  for c in "k256 22" "p256 22" "p384 20" "sm2 20"; do
    set -- $c
    [ "$v" != main ] && [ "$v" != w5 ] && [ "$v" != w6 ] && [ "$1" != k256 ] && [ "$1" != p256 ] && continue
    timeout 300 python scripts/prof_one.py $1 verify_keys $2 3 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s4_ab_keytab_variants.txt
unset ECB200_LIB
echo "== per-kernel times (ncu launch list) main, k256 + p256"
for c in k256 p256; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/s4_launches_$c.csv python scripts/prof_one.py $c verify_keys 22 1 > /dev/null 2>&1
  python - $O/s4_launches_$c.csv <<'PY'
import csv, sys
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    if len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    print("   %-46s %10.3f ms" % (d["Kernel Name"].split("(")[0][-46:], float(d["Metric Value"].replace(",", "")) / 1e6))
PY
done | tee $O/s4_launch_times.txt
echo "== round-2 GPU tests on the w5 and w6 libraries (keytab tests only)"
for v in w5 w6; do ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -k keytab 2>&1 | tail -2 | sed "s/^/$v /"; done | tee $O/s4_pytest_variants.txt
