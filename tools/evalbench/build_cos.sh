#!/bin/bash
# cos_pinned variants x pair protocol x launch bounds for the F_1_8D evaluation kernel (cross-compiled here, run on the GPU box)
cd "$(dirname "$0")"
mkdir -p _build
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
n=0
b() { # name flags...
  name=$1; shift
  nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 "$@" $common -o _build/cos_$name &
  n=$((n+1)); if [ $((n % 8)) = 0 ]; then wait; fi
}
b lib_mb7 -DPGN_GENZ_LIBRARY_MATH -DPGN_EVAL_MIN_BLOCKS=7
b lib_mb6 -DPGN_GENZ_LIBRARY_MATH -DPGN_EVAL_MIN_BLOCKS=6
for v in 1 2 3 4; do
  for mb in 6 7; do
    b v${v}p0_mb$mb -DPGN_COS_VARIANT=$v -DPGN_GENZ_NO_PAIR -DPGN_EVAL_MIN_BLOCKS=$mb
    b v${v}p1_mb$mb -DPGN_COS_VARIANT=$v -DPGN_EVAL_MIN_BLOCKS=$mb
  done
done
for v in 3 4; do for mb in 4 5 8; do b v${v}p1_mb$mb -DPGN_COS_VARIANT=$v -DPGN_EVAL_MIN_BLOCKS=$mb; done; done
wait
ls _build | grep -c cos_
