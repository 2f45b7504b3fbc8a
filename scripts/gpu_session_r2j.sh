#!/bin/bash
# Round-2 GPU session J: bench (both arms) on the build with 128-bit accesses / wider per-key windows / shuffle fetch, ncu --set full
# captures of every BASELINE config on that build (profiles/summary.json is what bench.py folds into `roofline`), ncu launch list
# of the bench command, dynamic CT audit of the shipped library
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
{ nvidia-smi -L; nproc; free -g | head -2; } > $O/s10_box.txt 2>&1
echo "== bench"
timeout 1200 python bench.py > $O/s10_bench.json 2> $O/s10_bench.err; echo "bench rc=$?"; cut -c1-500 $O/s10_bench.json; tail -3 $O/s10_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s10_bench_ref.json 2> $O/s10_bench_ref.err; echo "ref rc=$?"; cut -c1-300 $O/s10_bench_ref.json
timeout 600 python bench.py --curve p256 --no-others --no-cpu > $O/s10_bench_p256.json 2> $O/s10_bench_p256.err; echo "p256 rc=$?"; cut -c1-300 $O/s10_bench_p256.json
echo "== ncu captures"
cp profiles/summary.json $O/summary_r02j.json
export ECB200_SUMMARY_JSON=$PWD/$O/summary_r02j.json
cap() { # key curve op log2 rows title keep
  local key=$1 curve=$2 op=$3 lg=$4
  timeout 300 python scripts/prof_one.py $curve $op $lg 2 > $O/s10_prof_$key.txt 2>&1 || { echo "plain run failed: $key"; tail -5 $O/s10_prof_$key.txt; return; }
  tail -1 $O/s10_prof_$key.txt
  timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/r02j_$key python scripts/prof_one.py $curve $op $lg 1 > $O/s10_ncu_$key.log 2>&1
  python tools/ncu_op_summary.py $O/r02j_$key.ncu-rep $O/r02_ncu_$key.md $key $((1 << lg)) "$5, n = 2^$lg rows, final round-2 build" > $O/s10_sum_$key.txt 2>&1; tail -6 $O/s10_sum_$key.txt
  [ "$6" = keep ] || rm -f $O/r02j_$key.ncu-rep
}
cap verify_k256 k256 verify_keys 22 "ecb200_ecdsa_verify_dev secp256k1, 2^16 keys reused (BASELINE configs[2]): per-key tables" keep
cap verify_p256 p256 verify_keys 22 "ecb200_ecdsa_verify_dev P-256, 2^16 keys reused (BASELINE configs[3]): per-key tables"
cap mul_gen_k256 k256 mul_gen 16 "ecb200_mul_gen_dev secp256k1, FLAG_CT (BASELINE configs[0]): split fixed-base path, shuffle fetch"
cap verify_k256_rowpath k256 verify 22 "ecb200_ecdsa_verify_dev secp256k1, every key distinct: per-row path"
cap verify_p256_rowpath p256 verify 22 "ecb200_ecdsa_verify_dev P-256, every key distinct: per-row path"
cap mul_var_k256 k256 mul_var_proj 20 "ecb200_mul_var_dev secp256k1, X:Y:Z inputs, public scalars (BASELINE configs[1])"
cap mul_var_k256_ct k256 mul_var_proj_ct 20 "ecb200_mul_var_dev secp256k1, X:Y:Z inputs, constant-time (BASELINE configs[1])"
cap mul_var_p384 p384 mul_var 20 "ecb200_mul_var_dev P-384, public scalars (BASELINE configs[4])"
cap mul_var_sm2 sm2 mul_var 20 "ecb200_mul_var_dev SM2, public scalars (BASELINE configs[4])"
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/s10_ncu_bench.log 2>&1; echo "launch list rc=$?"
echo "== dynamic constant-time audit (shipped library)"
bash scripts/ct_audit.sh > $O/s10_ct_audit.log 2>&1; tail -5 $O/s10_ct_audit.log
ls -la $O | tail -30; du -sh $O
