#!/bin/bash
# Round-2 GPU session Q (final build): smoke(), whole GPU test tier, bench (both arms, P-256 headline), ncu launch list of the bench
# command, ncu --set full re-capture of the operations whose kernels changed since session J (verify on tables, fixed-base)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -2 ) | tee $O/s17_smoke.txt
echo "== bench"
timeout 1200 python bench.py > $O/s17_bench.json 2> $O/s17_bench.err; echo "bench rc=$?"; cut -c1-400 $O/s17_bench.json; tail -2 $O/s17_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s17_bench_ref.json 2> $O/s17_bench_ref.err; echo "ref rc=$?"; cut -c1-200 $O/s17_bench_ref.json
timeout 600 python bench.py --curve p256 --no-others --no-cpu > $O/s17_bench_p256.json 2> $O/s17_bench_p256.err; echo "p256 rc=$?"; cut -c1-200 $O/s17_bench_p256.json
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/s17_ncu_bench.log 2>&1; echo "launch list rc=$?"
echo "== ncu captures"
cp profiles/summary.json $O/summary_r02q.json
export ECB200_SUMMARY_JSON=$PWD/$O/summary_r02q.json
cap() { # key curve op log2 rows title
  local key=$1 curve=$2 op=$3 lg=$4
  timeout 300 python scripts/prof_one.py $curve $op $lg 2 > $O/s17_prof_$key.txt 2>&1 || { echo "plain run failed: $key"; tail -5 $O/s17_prof_$key.txt; return; }
  tail -1 $O/s17_prof_$key.txt
  timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/r02q_$key python scripts/prof_one.py $curve $op $lg 1 > $O/s17_ncu_$key.log 2>&1
  python tools/ncu_op_summary.py $O/r02q_$key.ncu-rep $O/r02_ncu_$key.md $key $((1 << lg)) "$5, n = 2^$lg rows, final round-2 build" > $O/s17_sum_$key.txt 2>&1; tail -4 $O/s17_sum_$key.txt
  rm -f $O/r02q_$key.ncu-rep
}
cap verify_k256 k256 verify_keys 22 "ecb200_ecdsa_verify_dev secp256k1, 2^16 keys reused (BASELINE configs[2]): per-key tables"
cap verify_p256 p256 verify_keys 22 "ecb200_ecdsa_verify_dev P-256, 2^16 keys reused (BASELINE configs[3]): per-key tables"
cap mul_gen_k256 k256 mul_gen 16 "ecb200_mul_gen_dev secp256k1, FLAG_CT (BASELINE configs[0]): split fixed-base path, shuffle fetch"
cap verify_k256_rowpath k256 verify 22 "ecb200_ecdsa_verify_dev secp256k1, every key distinct: per-row path"
echo "== tests"
( timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 ) | tee $O/s17_pytest.txt
