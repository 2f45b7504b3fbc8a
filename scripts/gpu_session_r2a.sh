#!/bin/bash
# Round-2 GPU session A (one gpurun call): GPU test tiers, bench (both arms), ncu captures of every BASELINE config on the shipped
# build (summarised on the box: the raw reports exceed gpurun's 64 MiB return limit), the dynamic constant-time audit, and the
# pointloop call-strategy variants.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{ nvidia-smi -L; nproc; free -g | head -2; } > $O/s1_box.txt 2>&1
echo "== tests"
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py > $O/s1_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/s1_pytest_gpu.log
( time timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --durations=10 ) > $O/s1_pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; tail -16 $O/s1_pytest_fullsize.log
echo "== bench"
timeout 900 python bench.py > $O/s1_bench.json 2> $O/s1_bench.err; echo "bench rc=$?"; cut -c1-600 $O/s1_bench.json; tail -3 $O/s1_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s1_bench_ref.json 2> $O/s1_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/s1_bench_ref.json
echo "== pointloop variants"
for v in nn_4 nn_6 nn_7 ni_3 ni_4 ni_5 ni_6 ii_3 ii_4; do echo "-- $v"; timeout 120 ./bench/pl2_$v 64; done > $O/s1_pointloop.txt 2>&1; cat $O/s1_pointloop.txt
echo "== ncu captures"
export ECB200_SUMMARY_JSON=$PWD/$O/summary_r02.json
cap() { # key curve op log2 rows title keep
  local key=$1 curve=$2 op=$3 lg=$4
  timeout 300 python scripts/prof_one.py $curve $op $lg 2 > $O/s1_prof_$key.txt 2>&1 || { echo "plain run failed: $key"; tail -5 $O/s1_prof_$key.txt; return; }
  cat $O/s1_prof_$key.txt
  timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/r02_$key python scripts/prof_one.py $curve $op $lg 1 > $O/s1_ncu_$key.log 2>&1
  python tools/ncu_op_summary.py $O/r02_$key.ncu-rep $O/r02_ncu_$key.md $key $((1 << lg)) "$5, n = 2^$lg rows, shipped round-2 build" > $O/s1_sum_$key.txt 2>&1; tail -6 $O/s1_sum_$key.txt
  [ "$6" = keep ] || rm -f $O/r02_$key.ncu-rep
}
cap verify_k256 k256 verify 22 "ecb200_ecdsa_verify_dev secp256k1 (BASELINE configs[2])" keep
cap mul_gen_k256 k256 mul_gen 16 "ecb200_mul_gen_dev secp256k1, constant-time (BASELINE configs[0])"
cap mul_var_k256 k256 mul_var_proj 20 "ecb200_mul_var_dev secp256k1, X:Y:Z inputs, public scalars (BASELINE configs[1])"
cap mul_var_k256_ct k256 mul_var_proj_ct 20 "ecb200_mul_var_dev secp256k1, X:Y:Z inputs, constant-time (BASELINE configs[1])"
cap verify_p256 p256 verify 22 "ecb200_ecdsa_verify_dev P-256 (BASELINE configs[3])"
cap mul_var_p384 p384 mul_var 20 "ecb200_mul_var_dev P-384, public scalars (BASELINE configs[4])"
cap mul_var_sm2 sm2 mul_var 20 "ecb200_mul_var_dev SM2, public scalars (BASELINE configs[4])"
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/s1_ncu_bench.log 2>&1; echo "launch list rc=$?"
echo "== dynamic constant-time audit"
bash scripts/ct_audit.sh > $O/s1_ct_audit.log 2>&1; tail -5 $O/s1_ct_audit.log
ls -la $O | tail -40; du -sh $O
