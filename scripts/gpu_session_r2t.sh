#!/bin/bash
# Round-2 GPU session T (the last three GPU-minutes): smoke() and the headline line of the other BASELINE operations with the
# re-captured profiles/summary.json
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
LIMIT=${1:-170}
( timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee $O/s20_smoke.txt; echo "smoke at ${SECONDS}s"
run() { # name args...
  local name=$1; shift
  local left=$((LIMIT - SECONDS))
  if [ $left -lt 35 ]; then echo "skipped $name: $left s left"; return; fi
  timeout $left python bench.py "$@" --no-others --no-cpu > $O/s20_bench_$name.json 2> $O/s20_bench_$name.err; echo "$name rc=$? at ${SECONDS}s"; cut -c1-160 $O/s20_bench_$name.json
}
run mul_var_k256 --op mul_var --curve k256
run mul_gen_k256 --op mul_gen --curve k256
run mul_var_p384 --op mul_var --curve p384
run mul_var_sm2 --op mul_var --curve sm2
echo "done at ${SECONDS}s"
