#!/bin/bash
# Round-2 GPU session G: divsteps inversion in the Montgomery-trick kernels (A/B against the Fermat build), shuffle-based table
# fetch of the split fixed-base kernel, k_kt_fill occupancy; parity tests of the new pieces; dynamic CT audit of the new kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 1200 python -m pytest tests/test_gpu_fixed_base_split.py tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 ) | tee $O/s7_pytest.txt
for v in main fermat g2shfl ktf6 ktf6e8; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so
  case $v in
    g2shfl) CASES=("k256 mul_gen 16" "k256 mul_gen 20" "p256 mul_gen 20" "p384 mul_gen 18") ;;
    ktf6|ktf6e8) CASES=("k256 verify_keys 22" "p256 verify_keys 22") ;;
    *) CASES=("k256 mul_gen 16" "k256 mul_gen 20" "p256 mul_gen 20" "p384 mul_gen 18" "k256 verify_keys 22" "p256 verify_keys 22" "k256 verify 22" "p256 verify 22" "k256 mul_var_proj 20" "p384 mul_var 20" "sm2 mul_var 20" "k256 sign 20" "p384 verify_keys 20") ;;
  esac
  for c in "${CASES[@]}"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s7_ab.txt
unset ECB200_LIB
echo "== per-kernel times (ncu launch list): k256 / p256 verify on per-key tables, main then fermat"
for v in main fermat; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so
  for c in k256 p256; do
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/s7_launches_${v}_$c.csv python scripts/prof_one.py $c verify_keys 22 1 > /dev/null 2>&1
    echo "-- $v $c"
    python - $O/s7_launches_${v}_$c.csv <<'PY'
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
  done
done | tee $O/s7_launch_times.txt
unset ECB200_LIB
echo "== dynamic CT audit (scan build = shipped), split fixed-base + signing"
bash scripts/ct_audit.sh k256:mul_gen k256:sign p256:mul_gen p384:mul_gen k256:mul_var > $O/s7_ct.log 2>&1
tail -30 $O/s7_ct.log
