#!/bin/bash
# round 2, session Y: F_1_8D evaluation kernel, warp-uniform warp index, per-lane weight look-up everywhere (session X's
# uniform half-chunk weights off): run-time runs or not x launch bound 4 / 5; plus the 5-D kernel (BASELINE config 1)
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2y_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -DPGN_EVAL_UNIFORM_WARP=1 -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=${4:-8} -DPGN_EVAL_MIN_BLOCKS=${3:-4} $2 $common -o _build/r2y_$1 & }
b runs4   "-DPGN_EVAL_RUNS=1" 4
b runs5   "-DPGN_EVAL_RUNS=1" 5
b noruns4 "" 4
b noruns5 "" 5
b noruns6 "" 6
b d5_noruns4 "" 4 5
b d5_noruns5 "" 5 5
b d5_runs4 "-DPGN_EVAL_RUNS=1" 4 5
wait
ls _build | grep -c r2y_
