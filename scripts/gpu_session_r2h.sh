#!/bin/bash
# Round-2 GPU session H: 128-bit accesses in the table kernels (k_kt_fill, k_wintab, normalise, prep), loads sunk below the window
# loops, wave-balanced k_kt_fill; A/B against the previous build; window widths re-measured with the cheaper fill; shuffle fetch of
# the split fixed-base kernel with its dynamic CT audit; whole GPU test tier on the new build
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 ) | tee $O/s8_pytest.txt
V=rustcrypto-elliptic-curves_b200/variants
for v in main prev g2shfl ktw6 ktw5 ept16; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/$V/libecb200_$v.so
  case $v in
    g2shfl) CASES=("k256 mul_gen 16" "k256 mul_gen 20" "p256 mul_gen 20" "p384 mul_gen 18" "k256 sign 20") ;;
    ktw6) CASES=("k256 verify_keys 22") ;;
    ktw5) CASES=("p256 verify_keys 22") ;;
    ept16) CASES=("k256 verify_keys 22" "p256 verify_keys 22") ;;
    *) CASES=("k256 verify_keys 22" "p256 verify_keys 22" "k256 verify 22" "p256 verify 22" "k256 mul_gen 16" "k256 mul_gen 20" "k256 mul_var_proj 20" "p384 mul_var 20" "sm2 mul_var 20" "k256 sign 20" "p384 verify_keys 20" "p384 verify 20") ;;
  esac
  for c in "${CASES[@]}"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s8_ab.txt
unset ECB200_LIB
echo "== per-kernel times (ncu launch list)"
for v in main prev; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/$V/libecb200_$v.so
  for c in "k256 verify_keys 22" "p256 verify_keys 22" "p256 verify 20"; do
    set -- $c
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/s8_launches_${v}_$1_$2.csv python scripts/prof_one.py $1 $2 $3 1 > /dev/null 2>&1
    echo "-- $v $c"
    python - $O/s8_launches_${v}_$1_$2.csv <<'PY'
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
done | tee $O/s8_launch_times.txt
echo "== dynamic CT audit of the shuffle build"
export ECB200_LIB=$PWD/$V/libecb200_g2shfl.so
bash scripts/ct_audit.sh k256:mul_gen k256:sign p256:mul_gen p384:mul_gen > $O/s8_ct.log 2>&1
cp $O/ct_audit_dynamic.md $O/s8_ct_audit_g2shfl.md
tail -25 $O/s8_ct.log
