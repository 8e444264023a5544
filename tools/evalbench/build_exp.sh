#!/bin/bash
# exp_pinned experiment (next step of DESIGN.md section 9, item 3): exp-based evaluation kernels, library exp vs pinned,
# launch bounds; also builds a second library with PGN_GENZ_FAST_EXP=1 for the parity run (tools/gpu_exp_check.sh)
cd "$(dirname "$0")"
mkdir -p _build
rm -f _build/exp_*
R=../..
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
for f in "CZero 8" "CZero 10" "Discontinuous 10" "Gaussian 8"; do
  set -- $f
  nvcc -I$R/include -DBENCH_FAMILY=$1 -DBENCH_NDIM=$2 -DPGN_EVAL_MIN_BLOCKS=4 $common -o _build/exp_$1$2_lib_mb4 &
  for mb in 3 4 6; do
    nvcc -I$R/include -DBENCH_FAMILY=$1 -DBENCH_NDIM=$2 -DPGN_GENZ_FAST_EXP=1 -DPGN_EVAL_MIN_BLOCKS=$mb $common -o _build/exp_$1$2_fast_mb$mb &
  done
  wait
done
mkdir -p _build/fastexp
make -C $R/gpuintegration_b200/csrc -j8 EXTRA=-DPGN_GENZ_FAST_EXP=1 OUT=../../tools/evalbench/_build/fastexp/libpagani_b200.so OBJDIR=build_fastexp 2>&1 | grep -E "error|Error"
ls _build | grep -c exp_
