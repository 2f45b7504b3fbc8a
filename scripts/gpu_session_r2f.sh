#!/bin/bash
# Round-2 GPU session F: split fixed-base path (k_gen_half + k_sum_normalize) - parity tests, A/B against the one-thread
# kernel, window width / resident-CTA variants, per-kernel times; keytab tests with the per-curve window width
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_fixed_base_split.py tests/test_libcrypto_cross.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 ) | tee $O/s6_pytest_split.txt
( timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_next_rows.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 ) | tee $O/s6_pytest_round2.txt
for v in main one g2w4 g2c6 g2c8; do
  unset ECB200_LIB ECB200_GEN2
  case $v in
    main) ;;
    one) export ECB200_GEN2=0 ;;
    *) export ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so ;;
  esac
  for c in "k256 mul_gen 16" "k256 mul_gen 18" "k256 mul_gen 20" "k256 sign 20" "p256 mul_gen 16" "p256 mul_gen 20" "p384 mul_gen 18" "sm2 mul_gen 20"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s6_ab_gen2.txt
unset ECB200_LIB ECB200_GEN2
echo "== per-kernel times (ncu launch list), k256 mul_gen 2^16: split path, then one-thread path"
for g in 1 0; do
  ECB200_GEN2=$g timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/s6_launches_gen$g.csv python scripts/prof_one.py k256 mul_gen 16 1 > /dev/null 2>&1
  python - $O/s6_launches_gen$g.csv <<'PY'
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
done | tee $O/s6_launch_times.txt
echo "== ncu --set full of the split path, k256 mul_gen 2^16"
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -o $O/r02_mul_gen_k256_split -f python scripts/prof_one.py k256 mul_gen 16 1 > $O/s6_ncu.log 2>&1
tail -2 $O/s6_ncu.log
