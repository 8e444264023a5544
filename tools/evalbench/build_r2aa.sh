#!/bin/bash
# round 2, session AA: every Genz family with its shipped tuning, warp-uniform warp index off / on; the unit-Jacobian
# kernel variant off / on for the two 10-D families that had it switched on unmeasured; 5-D oscillatory bounds
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2aa_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=$2 -DBENCH_NDIM=$3 -DPGN_EVAL_MIN_BLOCKS=$4 $5 $common -o _build/r2aa_$1 & }
for uw in 0 1; do
b osc8_uw$uw   Oscillatory 8 5 "-DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -DPGN_EVAL_UNIFORM_WARP=$uw"
b osc5_mb5_uw$uw Oscillatory 5 5 "-DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -DPGN_EVAL_UNIFORM_WARP=$uw"
b osc5_mb4_uw$uw Oscillatory 5 4 "-DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -DPGN_EVAL_UNIFORM_WARP=$uw"
for uj in 0 1; do
b c010_uj${uj}_uw$uw CZero 10 4 "-DPGN_EVAL_SEGMENTS=1 -DPGN_EVAL_UNIT_JACOBIAN=$uj -DPGN_EVAL_UNIFORM_WARP=$uw"
b disc10_uj${uj}_uw$uw Discontinuous 10 4 "-DPGN_EVAL_UNIT_JACOBIAN=$uj -DPGN_EVAL_UNIFORM_WARP=$uw"
done
b gauss8_uw$uw Gaussian 8 3 "-DPGN_EVAL_UNIFORM_WARP=$uw"
b gauss5_uw$uw Gaussian 5 3 "-DPGN_EVAL_UNIFORM_WARP=$uw"
b pp6_uw$uw ProductPeak 6 4 "-DPGN_EVAL_UNIFORM_WARP=$uw"
b cp8_uw$uw CornerPeak 8 8 "-DPGN_EVAL_UNIFORM_WARP=$uw"
wait
done
ls _build | grep -c r2aa_
