#!/bin/bash
# round 2, session X: F_1_8D evaluation kernel with a warp-uniform warp index: rule weights of single-set half-chunks
# of the mixed trips as constant-bank operands behind a uniform branch (per-lane shared-memory reads only for the
# half-chunk that straddles the boundary), with / without run-time runs, launch bounds 4 / 5 / 6
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2x_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -DPGN_EVAL_UNIFORM_WARP=1 -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=${3:-4} $2 $common -o _build/r2x_$1 & }
b runs4   "-DPGN_EVAL_RUNS=1" 4
b runs5   "-DPGN_EVAL_RUNS=1" 5
b noruns4 "" 4
b noruns5 "" 5
b noruns6 "" 6
wait
ls _build | grep -c r2x_
