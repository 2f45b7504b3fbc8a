#!/bin/bash
# Round-2 GPU session P: L2 prefetch of the next (key, window) item in k_kt_fill (main vs nopf) and of the next window's table
# entries in k_verify_keytab (pfmain)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
V=rustcrypto-elliptic-curves_b200/variants
for v in main nopf pfmain main nopf pfmain; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/$V/libecb200_$v.so
  for c in "k256 verify_keys 22" "p256 verify_keys 22" "p384 verify_keys 20"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s16_ab_prefetch.txt
for v in main nopf; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/$V/libecb200_$v.so
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/s16_launches_${v}.csv python scripts/prof_one.py k256 verify_keys 22 1 > /dev/null 2>&1
  echo "-- $v k256 verify_keys 22"
  python - $O/s16_launches_${v}.csv <<'PY'
import csv, sys
hdr = None
for r in csv.reader(open(sys.argv[1])):
    if hdr is None:
        if "Kernel Name" in r: hdr = r
        continue
    if len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    print("   %-46s %10.4f ms" % (d["Kernel Name"].split("(")[0][-46:], float(d["Metric Value"].replace(",", "")) / 1e6))
PY
done | tee -a $O/s16_ab_prefetch.txt
