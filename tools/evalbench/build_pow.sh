#!/bin/bash
# pow2_pinned experiment: Gaussian<8> / ProductPeak<6> evaluation kernels, library pow vs pinned, launch bounds
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/pow_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
for f in "Gaussian 8" "ProductPeak 6"; do
  set -- $f
  nvcc -I$R/include -DBENCH_FAMILY=$1 -DBENCH_NDIM=$2 -DPGN_EVAL_MIN_BLOCKS=8 $common -o _build/pow_$1$2_lib_mb8 &
  for mb in 3 4 5 6 8; do
    nvcc -I$R/include -DBENCH_FAMILY=$1 -DBENCH_NDIM=$2 -DPGN_GENZ_FAST_POW=1 -DPGN_EVAL_MIN_BLOCKS=$mb $common -o _build/pow_$1$2_fast_mb$mb &
  done
  wait
done
ls _build | grep -c pow_
