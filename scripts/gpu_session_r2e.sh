#!/bin/bash
# Round-2 GPU session E: GPU test tier on the restored tree, then session D (per-key table window widths / resident CTAs)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > $O/s5_pytest_gpu.txt 2>&1
tail -3 $O/s5_pytest_gpu.txt
bash scripts/gpu_session_r2d.sh
