#!/bin/bash
# launch-bound sweep of the evaluation kernel (cross-compiled on the CPU box, run on the GPU box)
cd "$(dirname "$0")"
mkdir -p _build
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
for mb in 4 6 7 8; do
  for f in "Oscillatory 8" "CZero 8" "Gaussian 8" "CZero 10"; do
    set -- $f
    nvcc -I$R/include -DBENCH_FAMILY=$1 -DBENCH_NDIM=$2 -DPGN_EVAL_MIN_BLOCKS=$mb $common -o _build/eb_$1$2_mb$mb &
  done
  nvcc -I$R/include -DBENCH_FAMILY=Product -DBENCH_TESTFN -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=$mb $common -o _build/eb_Product8_mb$mb &
  wait
done
for mb in 3 4; do
  nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=$mb -DPGN_EVAL_WARPS=8 $common -o _build/eb_Oscillatory8_w8mb$mb &
done
wait
ls _build | wc -l
