#!/bin/bash
# round 2, session W: F_1_8D evaluation kernel on top of session V's best (unit Jacobian + range promise + runs):
# parity test of the cosine as one logic instruction with a predicate result, 64-bit shuffles low word first,
# warp index broadcast from lane 0 (warp-uniform for the compiler: no divergent twins of the shuffle groups)
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2w_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 -DPGN_EVAL_MIN_BLOCKS=${3:-4} $2 $common -o _build/r2w_$1 & }
b v_both   "-DPGN_EVAL_RUNS=1 -DPGN_EVAL_SHFL_LO_FIRST=0"
b odd2     "-DPGN_EVAL_RUNS=1 -DPGN_EVAL_SHFL_LO_FIRST=0 -DPGN_COS_ODD_FORM=2"
b odd2lo   "-DPGN_EVAL_RUNS=1 -DPGN_COS_ODD_FORM=2"
b uw       "-DPGN_EVAL_RUNS=1 -DPGN_EVAL_UNIFORM_WARP=1"
b uwodd2   "-DPGN_EVAL_RUNS=1 -DPGN_EVAL_UNIFORM_WARP=1 -DPGN_COS_ODD_FORM=2"
b uwodd2_5 "-DPGN_EVAL_RUNS=1 -DPGN_EVAL_UNIFORM_WARP=1 -DPGN_COS_ODD_FORM=2" 5
b uwnoruns "-DPGN_EVAL_UNIFORM_WARP=1 -DPGN_COS_ODD_FORM=2"
b uwnoruns5 "-DPGN_EVAL_UNIFORM_WARP=1 -DPGN_COS_ODD_FORM=2" 5
b uwchk    "-DPGN_EVAL_RUNS=1 -DPGN_EVAL_UNIFORM_WARP=1 -DPGN_COS_ODD_FORM=2 -DPGN_EVAL_NOCHECK=0"
wait
ls _build | grep -c r2w_
