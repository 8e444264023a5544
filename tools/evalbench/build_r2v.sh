#!/bin/bash
# round 2, session V: F_1_8D evaluation kernel -- unit-Jacobian variant (base) plus: launch-time range promise
# (no per-point |s| < 2^31 checks), middle trips as run-time runs, unrolled middle loop
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2v_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=${3:-4} $2 $common -o _build/r2v_$1 & }
b base   "-DPGN_EVAL_NOCHECK=0"
b nochk  "-DPGN_EVAL_NOCHECK=1"
b runs   "-DPGN_EVAL_NOCHECK=0 -DPGN_EVAL_RUNS=1"
b both   "-DPGN_EVAL_NOCHECK=1 -DPGN_EVAL_RUNS=1"
b u2     "-DPGN_EVAL_NOCHECK=0 -DPGN_EVAL_MIDDLE_UNROLL=2"
b u2nc   "-DPGN_EVAL_NOCHECK=1 -DPGN_EVAL_MIDDLE_UNROLL=2"
b nochk5 "-DPGN_EVAL_NOCHECK=1" 5
b both5  "-DPGN_EVAL_NOCHECK=1 -DPGN_EVAL_RUNS=1" 5
b runs5  "-DPGN_EVAL_NOCHECK=0 -DPGN_EVAL_RUNS=1" 5
b u2nc5  "-DPGN_EVAL_NOCHECK=1 -DPGN_EVAL_MIDDLE_UNROLL=2" 5
wait
ls _build | grep -c r2v_
