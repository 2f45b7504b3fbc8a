#!/bin/bash
# Round-2 GPU session O: mixed additions ordered for register pressure (outputs stored as soon as their inputs are dead): A/B
# against the build before, whole GPU test tier, dynamic CT audit of the fixed-base path (madd_ct changed)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
V=rustcrypto-elliptic-curves_b200/variants
for v in main prev2; do
  unset ECB200_LIB
  [ "$v" != main ] && export ECB200_LIB=$PWD/$V/libecb200_$v.so
  for c in "k256 verify_keys 22" "p256 verify_keys 22" "k256 verify 22" "p256 verify 22" "k256 mul_gen 16" "k256 mul_gen 20" "k256 sign 20" "k256 mul_var_proj 20" "p384 mul_var 20" "sm2 mul_var 20" "p384 verify_keys 20"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/$v /"
  done
done | tee $O/s15_ab_madd_order.txt
unset ECB200_LIB
( timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -3 ) | tee $O/s15_pytest.txt
bash scripts/ct_audit.sh k256:mul_gen k256:sign p256:mul_gen > $O/s15_ct.log 2>&1
cp $O/ct_audit_dynamic.md $O/s15_ct_audit.md; tail -4 $O/s15_ct.log
