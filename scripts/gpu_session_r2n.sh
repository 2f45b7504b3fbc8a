#!/bin/bash
# Round-2 GPU session N: smoke() with the table path, the retuned table policy (tests), bench headline with the new clock
# sampler, occupancy variants of k_kt_fill (5 / 6 resident CTAs) and k_verify_keytab (7 / 8) now that both are compute-bound
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3 ) | tee $O/s14_smoke.txt
( timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_fixed_base_split.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 ) | tee $O/s14_pytest.txt
timeout 600 python bench.py --no-others --no-cpu > $O/s14_bench_headline.json 2> $O/s14_bench.err; echo "bench rc=$?"; cut -c1-1200 $O/s14_bench_headline.json; tail -2 $O/s14_bench.err
V=rustcrypto-elliptic-curves_b200/variants
for v in main ktf5 ktf6 kt7 kt8; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/$V/libecb200_$v.so
  for c in "k256 verify_keys 22" "p256 verify_keys 22" "p384 verify_keys 20"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s14_ab_occupancy.txt
unset ECB200_LIB
