#!/bin/bash
# Dynamic constant-time audit (run under gpurun): the secret-scalar kernels are launched with the same public inputs and four
# different SECRET patterns; ncu's executed-instruction, divergence and memory-sector counters of every launch must be identical
# across the patterns (scripts/ct_audit_compare.py).  Output: gpurun_out/ct_<curve>_<op>_<pattern>.csv
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__sass_branch_targets_threads_divergent.sum,smsp__sass_branch_targets.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,smsp__inst_executed_op_shared_ld.sum
CASES=("$@")
[ ${#CASES[@]} -eq 0 ] && CASES=(k256:mul_var k256:mul_gen k256:sign p256:mul_var p256:sign p384:mul_gen)
for case in "${CASES[@]}"; do
  curve=${case%%:*}; op=${case##*:}
  timeout 300 python scripts/ct_dyn_run.py $curve $op > gpurun_out/ct_plain.log 2>&1 || { echo "plain run failed: $curve $op"; tail -5 gpurun_out/ct_plain.log; continue; }
  timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/ct_${curve}_${op}.csv \
    python scripts/ct_dyn_run.py $curve $op > /dev/null 2>&1
done
python scripts/ct_audit_compare.py gpurun_out | tee gpurun_out/ct_audit_dynamic.md
