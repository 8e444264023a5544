#!/bin/bash
# second sweep: cos variant {3,4} x four-point trips {0,1} x launch bounds, F_1_8D (cross-compiled here, run on the GPU box)
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/cos_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
n=0
b() { # name flags...
  name=$1; shift
  nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 "$@" $common -o _build/cos_$name &
  n=$((n+1)); if [ $((n % 8)) = 0 ]; then wait; fi
}
b lib_mb7 -DPGN_GENZ_LIBRARY_MATH -DPGN_EVAL_MIN_BLOCKS=7
for v in 3 4; do
  for q in 0 1; do
    for mb in 3 4 5 6 7; do
      b v${v}q${q}_mb$mb -DPGN_COS_VARIANT=$v -DPGN_EVAL_QUAD=$q -DPGN_EVAL_MIN_BLOCKS=$mb
    done
  done
done
b v4q1_w8mb2 -DPGN_COS_VARIANT=4 -DPGN_EVAL_QUAD=1 -DPGN_EVAL_WARPS=8 -DPGN_EVAL_MIN_BLOCKS=2
b v4q1_w8mb3 -DPGN_COS_VARIANT=4 -DPGN_EVAL_QUAD=1 -DPGN_EVAL_WARPS=8 -DPGN_EVAL_MIN_BLOCKS=3
b v4q1_w2mb8 -DPGN_COS_VARIANT=4 -DPGN_EVAL_QUAD=1 -DPGN_EVAL_WARPS=2 -DPGN_EVAL_MIN_BLOCKS=8
wait
ls _build | grep -c cos_
