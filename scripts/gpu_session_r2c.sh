#!/bin/bash
# Round-2 GPU session C: per-key window tables - GPU test tiers, quick A/B (tables vs per-row), bench, ncu of the new main kernel
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== tests (round-2 file first)"
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -p no:cacheprovider > $O/s3_pytest_round2.log 2>&1; echo "round2 rc=$?"; tail -4 $O/s3_pytest_round2.log
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py --deselect tests/test_gpu_round2.py > $O/s3_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -4 $O/s3_pytest_gpu.log
echo "== A/B at 2^22 rows, 2^16 keys"
for kt in 1 0; do
  for c in k256 p256; do ECB200_KEYTAB=$kt timeout 300 python scripts/prof_one.py $c verify_keys 22 3 2>&1 | tail -1 | sed "s/^/keytab=$kt /"; done
done | tee $O/s3_ab_keytab.txt
for c in p384 sm2; do timeout 300 python scripts/prof_one.py $c verify_keys 20 2 2>&1 | tail -1; done | tee -a $O/s3_ab_keytab.txt
echo "== bench"
timeout 1200 python bench.py > $O/s3_bench.json 2> $O/s3_bench.err; echo "bench rc=$?"; cut -c1-400 $O/s3_bench.json; tail -3 $O/s3_bench.err
echo "== ncu: the table path"
export ECB200_SUMMARY_JSON=$PWD/$O/summary_r02c.json
cap() {
  local key=$1 curve=$2 op=$3 lg=$4
  timeout 300 python scripts/prof_one.py $curve $op $lg 2 > $O/s3_prof_$key.txt 2>&1 || { echo "plain run failed: $key"; tail -5 $O/s3_prof_$key.txt; return; }
  timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/r02_$key python scripts/prof_one.py $curve $op $lg 1 > $O/s3_ncu_$key.log 2>&1
  python tools/ncu_op_summary.py $O/r02_$key.ncu-rep $O/r02_ncu_$key.md $key $((1 << lg)) "$5, n = 2^$lg rows" > $O/s3_sum_$key.txt 2>&1; tail -6 $O/s3_sum_$key.txt
  [ "$6" = keep ] || rm -f $O/r02_$key.ncu-rep
}
cap verify_k256 k256 verify_keys 22 "ecb200_ecdsa_verify_dev secp256k1, 2^16 keys reused (BASELINE configs[2]): per-key tables" keep
cap verify_p256 p256 verify_keys 22 "ecb200_ecdsa_verify_dev P-256, 2^16 keys reused (BASELINE configs[3]): per-key tables"
ls -la $O | tail -12
