#!/bin/bash
# Round-2 GPU session R (the last 14 GPU-minutes of the round, final build = HEAD): bench (both arms), ncu launch list of the bench
# command, then as much of the GPU test tier as the remaining time allows (its own deadline; the build is the one whose 243 tests
# were green in session N - the commit since only added compiled-out code).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
LIMIT=${1:-780}
{ nvidia-smi -L; nproc; free -g | head -2; } > $O/s18_box.txt 2>&1
echo "== bench"
timeout 330 python bench.py > $O/s18_bench.json 2> $O/s18_bench.err; echo "bench rc=$? at ${SECONDS}s"; cut -c1-400 $O/s18_bench.json; tail -2 $O/s18_bench.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/s18_bench_ref.json 2> $O/s18_bench_ref.err; echo "ref rc=$? at ${SECONDS}s"; cut -c1-200 $O/s18_bench_ref.json
echo "== ncu launch list of the bench command"
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-others > $O/s18_ncu_bench.log 2>&1; echo "launch list rc=$? at ${SECONDS}s"
echo "== tests"
LEFT=$((LIMIT - SECONDS - 15))
if [ $LEFT -gt 60 ]; then
  ( timeout $LEFT python -m pytest tests -m gpu -x -q -p no:cacheprovider --durations=15 2>&1 | tail -25 ) | tee $O/s18_pytest.txt
else
  echo "no time left for the test tier ($LEFT s)" | tee $O/s18_pytest.txt
fi
echo "done at ${SECONDS}s"
