#!/bin/bash
# Round-2 GPU session I: the new defaults (6-bit / 5-bit per-key windows, shuffle fetch) through the GPU test tier; width of the
# big fixed-base table (ECB200_GW: 16 = 34 MiB, 20 = 410 MB, 22 = 1.5 GB per curve) on the verify paths
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -5 ) | tee $O/s9_pytest.txt
for gw in 16 20 22 18; do
  export ECB200_GW=$gw
  for c in "k256 verify_keys 22" "p256 verify_keys 22" "k256 verify 22" "p384 verify_keys 20"; do
    set -- $c
    timeout 300 python scripts/prof_one.py $1 $2 $3 5 2>&1 | tail -1 | sed "s/^/gw=$gw /"
  done
done | tee $O/s9_ab_gw.txt
unset ECB200_GW
# correctness of a non-default width through the C ABI: the verify tests again at gw = 20 (straddling windows)
( ECB200_GW=20 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -x -q -p no:cacheprovider -k "verify or wycheproof or keytab" 2>&1 | tail -3 ) | tee $O/s9_pytest_gw20.txt
