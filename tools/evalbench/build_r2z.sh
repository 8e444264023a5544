#!/bin/bash
# round 2, session Z: F_1_8D evaluation kernel, warp-uniform warp index: {runs, no runs} x {launch bound 4, 5} x
# {parity test of the cosine: compare (0), predicate result of one logic instruction (2)}
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2z_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -DPGN_EVAL_LANE_WEIGHTS_SMEM=1 -DPGN_EVAL_UNIT_JACOBIAN=1 -DPGN_EVAL_UNIFORM_WARP=1 -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=${4:-8} -DPGN_EVAL_MIN_BLOCKS=${3:-4} $2 $common -o _build/r2z_$1 & }
for odd in 0 2; do for mb in 4 5; do
b runs${mb}_o$odd   "-DPGN_EVAL_RUNS=1 -DPGN_COS_ODD_FORM=$odd" $mb
b noruns${mb}_o$odd "-DPGN_COS_ODD_FORM=$odd" $mb
done; done
wait
ls _build | grep -c r2z_
