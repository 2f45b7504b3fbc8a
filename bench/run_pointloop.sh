#!/bin/bash
# builds the call-strategy variants of bench/pointloop.cu (here or on the GPU box) and runs them
set -e
cd "$(dirname "$0")"
NV="nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr"
NOI='__device__ __noinline__'
INL='__device__ __forceinline__'
build() { # name pointfn fieldfn extra
  [ -x pointloop_$1 ] || $NV -DECB_POINT_FN="$2" -DECB_FIELD_FN="$3" $4 -o pointloop_$1 pointloop.cu
}
build pn_fn "$NOI" "$NOI" ""
build pi_fn "$INL" "$NOI" ""
build pi_fi "$INL" "$INL" ""
build pn_fi "$NOI" "$INL" ""
# field functions that store their result through a pointer (operands still by value)
[ -x pointloop_pn_fo ] || $NV -DECB_POINT_FN="$NOI" -DECB_FIELD_OUT_PTR -o pointloop_pn_fo pointloop.cu
if [ "$1" = run ]; then
  for v in ${2:-pn_fn pi_fn pi_fi pn_fi pn_fo}; do echo "== $v"; ./pointloop_$v 64; done
fi
