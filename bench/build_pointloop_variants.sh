#!/bin/bash
# builds bench/pointloop.cu in the call-strategy x resident-CTA variants measured in round 2 (binaries travel to the GPU box)
# naming: pl2_<point fn: n = noinline, i = inline><field fn: n / i>_<min CTAs per SM>
cd "$(dirname "$0")"
NV="nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr"
NOI='__device__ __noinline__'
INL='__device__ __forceinline__'
b() { # name point field minctas
  $NV -DMINCTAS=$4 -DECB_POINT_FN="$2" -DECB_FIELD_FN="$3" $5 -o pl2_$1_$4 pointloop.cu &
}
for m in 4 6 7; do b nn "$NOI" "$NOI" $m; done
for m in 3 4 5 6; do b ni "$NOI" "$INL" $m; done
for m in 3 4; do b ii "$INL" "$INL" $m; done
wait
ls -la pl2_*
