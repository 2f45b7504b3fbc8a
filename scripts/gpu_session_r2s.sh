#!/bin/bash
# Round-2 GPU session S (the last GPU-minutes of the round, final build): bench with the lincomb2 output check, P-256 headline,
# ncu --set full re-capture of the verify call on per-key tables (kernels changed since session J: mixed-addition order)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
LIMIT=${1:-420}
echo "== bench"
timeout 150 python bench.py > $O/s19_bench.json 2> $O/s19_bench.err; echo "bench rc=$? at ${SECONDS}s"; cut -c1-300 $O/s19_bench.json; tail -2 $O/s19_bench.err
timeout 90 python bench.py --curve p256 --no-others --no-cpu > $O/s19_bench_p256.json 2> $O/s19_bench_p256.err; echo "p256 rc=$? at ${SECONDS}s"; cut -c1-200 $O/s19_bench_p256.json
echo "== ncu captures"
cp profiles/summary.json $O/summary_r02s.json
export ECB200_SUMMARY_JSON=$PWD/$O/summary_r02s.json
cap() { # key curve op log2 rows title
  local key=$1 curve=$2 op=$3 lg=$4
  local left=$((LIMIT - SECONDS))
  if [ $left -lt 110 ]; then echo "skipped $key: $left s left"; return; fi
  timeout $((left - 20)) ncu --set full --import-source on --clock-control none --profile-from-start off -f -o $O/r02s_$key python scripts/prof_one.py $curve $op $lg 1 > $O/s19_ncu_$key.log 2>&1; echo "ncu $key rc=$? at ${SECONDS}s"
  timeout 60 python tools/ncu_op_summary.py $O/r02s_$key.ncu-rep $O/r02_ncu_$key.md $key $((1 << lg)) "$5, n = 2^$lg rows, final round-2 build (session S)" > $O/s19_sum_$key.txt 2>&1; tail -4 $O/s19_sum_$key.txt
  rm -f $O/r02s_$key.ncu-rep
}
cap verify_k256 k256 verify_keys 22 "ecb200_ecdsa_verify_dev secp256k1, 2^16 keys reused (BASELINE configs[2]): per-key tables"
cap verify_p256 p256 verify_keys 22 "ecb200_ecdsa_verify_dev P-256, 2^16 keys reused (BASELINE configs[3]): per-key tables"
echo "done at ${SECONDS}s"
