#!/bin/bash
# Round-2 GPU session B: MOV-patch experiment on the pointloop cubins (driver API), A/B of the library with and without the
# SASS pass, GPU test tiers on the fixed build, dynamic CT audit
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
echo "== pointloop: ptxas cubin vs MOV-patched cubin"
for m in 4 6 7; do timeout 120 ./bench/pointloop_drv bench/pl_nn$m.cubin bench/pl_nn${m}_mov.cubin --minctas $m; done 2>&1 | tee $O/s2_pointloop_mov.txt
timeout 120 ./bench/pointloop_drv bench/pl_ii4.cubin bench/pl_ii4_mov.cubin --minctas 4 2>&1 | tee -a $O/s2_pointloop_mov.txt
echo "== library A/B: shipped vs MOV-patched"
for v in main mov; do
  if [ "$v" = main ]; then unset ECB200_LIB; else export ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_$v.so; fi
  echo "-- $v"; timeout 300 python scripts/quick_bench.py k256,p256 22 2>&1 | grep -E "verify|mul_var|mul_gen"
  timeout 300 python scripts/quick_bench.py p384,sm2 20 2>&1 | grep -E "verify|mul_var ct=0"
done 2>&1 | tee $O/s2_ab_mov.txt
echo "== tests on the MOV-patched library"
ECB200_LIB=$PWD/rustcrypto-elliptic-curves_b200/variants/libecb200_mov.so timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py -x > $O/s2_pytest_gpu_mov.log 2>&1; echo "pytest gpu (mov) rc=$?"; tail -5 $O/s2_pytest_gpu_mov.log
unset ECB200_LIB
echo "== tests on the shipped library"
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py > $O/s2_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 $O/s2_pytest_gpu.log
( time timeout 1500 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider --durations=10 ) > $O/s2_pytest_fullsize.log 2>&1; echo "pytest fullsize rc=$?"; tail -14 $O/s2_pytest_fullsize.log
echo "== dynamic constant-time audit"
bash scripts/ct_audit.sh > $O/s2_ct_audit.log 2>&1; tail -3 $O/s2_ct_audit.log
