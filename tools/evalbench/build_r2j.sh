#!/bin/bash
# round 2, session J: F_1_8D evaluation kernel, launch bound x bounds prefetch (tail-A-only off, no trip segments):
# how many of the cosine literals stay in registers across the trip loop depends on the register budget
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2j_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
for mb in 4 5; do
  nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=$mb $common -o _build/r2j_osc8_mb$mb &
  nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=$mb -DPGN_EVAL_PREFETCH_BOUNDS=0 $common -o _build/r2j_osc8_mb${mb}_np &
done
nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=5 -DPGN_EVAL_TAIL_A_ONLY=1 $common -o _build/r2j_osc8_mb5_tailA &
wait
ls _build | grep -c r2j_
