#!/bin/bash
# builds eval_bench_<family><ndim>_mb<k> for a few functors / launch-bound settings (CPU box, cross-compile)
set -e
cd "$(dirname "$0")"
mkdir -p _build
ROOT=../..
build() { # family ndim minblocks extra
  nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w \
    -I$ROOT/include -DBENCH_FAMILY=$1 -DBENCH_NDIM=$2 -DPGN_EVAL_MIN_BLOCKS=$3 $4 eval_bench.cu \
    -o _build/eb_$1$2_mb$3$5 -L$ROOT/gpuintegration_b200 -lpagani_b200 \
    -Xlinker -rpath -Xlinker '$ORIGIN/../../../gpuintegration_b200' &
}
for mb in 1 8; do
  build Oscillatory 8 $mb
  build Gaussian 8 $mb
  build CZero 8 $mb
  build ProductPeak 6 $mb
  build CZero 10 $mb
done
wait
ls -la _build
