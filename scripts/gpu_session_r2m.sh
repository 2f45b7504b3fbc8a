#!/bin/bash
# Round-2 GPU session M: whole GPU test tier on the build with per-call table widths and threaded staging copies; bench headline
# only (pageable-buffer leg); A/B of the narrow / wide tables at 8, 16, 32 and 64 rows per key
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -4 ) | tee $O/s13_pytest.txt
timeout 600 python bench.py --no-others --no-cpu > $O/s13_bench_headline.json 2> $O/s13_bench.err; echo "bench rc=$?"; cut -c1-900 $O/s13_bench_headline.json
for lg in 19 20 21 22; do
  for wide in 0 1; do
    for c in k256 p256; do
      ECB200_KT_WIDE=$wide timeout 300 python scripts/prof_one.py $c verify_keys $lg 5 2>&1 | tail -1 | sed "s/^/rows_per_key=$((1 << (lg - 16))) wide=$wide /"
    done
  done
done | tee $O/s13_ab_table_width_by_reuse.txt
