#!/bin/bash
# runs bench/imad_peak with the SM clock sampled alongside (nvidia-smi, 50 ms period)
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 50 > gpurun_out/imad_peak_clocks.csv &
SMI=$!
./bench/imad_peak > gpurun_out/peaks_int_v2.json
kill $SMI
sort gpurun_out/imad_peak_clocks.csv | uniq -c | sort -rn | head -5
cat gpurun_out/peaks_int_v2.json
