#!/bin/bash
# diagnostic variants of the evaluation kernel (timing experiments)
cd "$(dirname "$0")"
mkdir -p _build
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
nvcc -I$R/include -DBENCH_FAMILY=Product -DBENCH_TESTFN -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=1 $common -o _build/eb_Product8_mb1 &
[ -d /tmp/old_inc ] && nvcc -I/tmp/old_inc -I$R/include -DBENCH_FAMILY=Product -DBENCH_TESTFN -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=0 $common -o _build/eb_Product8_old &
nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=1 -DPGN_DEBUG_UNIT_WEIGHTS $common -o _build/eb_Oscillatory8_unitw &
nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=4 -DPGN_EVAL_WARPS=8 $common -o _build/eb_Oscillatory8_w8 &
nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=16 -DPGN_EVAL_WARPS=2 $common -o _build/eb_Oscillatory8_w2 &
nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=1 $common -o _build/eb_Oscillatory8_mb1 &
wait
ls _build
