#!/bin/bash
# round 2, session AC: coordinate table rows of the evaluation kernel padded from 11 to 16 doubles (one 128-byte line
# per axis; knob PGN_EVAL_X_STRIDE, removed again after this session: no effect): does a 64-bit gather whose 32 addresses lie in ONE aligned 128-byte window cost one wavefront instead of two?
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/r2ac_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
b() { nvcc -I$R/include -DBENCH_FAMILY=$2 -DBENCH_NDIM=$3 $4 $common -o _build/r2ac_$1 & }
for xs in 11 12 16; do
b osc8_x$xs Oscillatory 8 "-DPGN_EVAL_X_STRIDE=$xs"
b osc5_x$xs Oscillatory 5 "-DPGN_EVAL_X_STRIDE=$xs"
b c010_x$xs CZero 10 "-DPGN_EVAL_X_STRIDE=$xs"
b disc10_x$xs Discontinuous 10 "-DPGN_EVAL_X_STRIDE=$xs"
b cp8_x$xs CornerPeak 8 "-DPGN_EVAL_X_STRIDE=$xs"
done
wait
ls _build | grep -c r2ac_
