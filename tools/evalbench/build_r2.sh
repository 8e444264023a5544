#!/bin/bash
# round 2: evaluation-kernel variants (cross-compiled here, run on the GPU box by tools/gpu_sessions/r2_*.sh)
#   old_*   the round-1 kernel (include/ tree of commit 55affdb), new_* the working tree
cd "$(dirname "$0")"
mkdir -p _build/old_include
rm -f _build/r2_*
R=../..
(cd $R && git archive 55affdb include) | tar -x -C _build/old_include --strip-components=1
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
n=0
b() { # name include-dir family ndim flags...
  name=$1; inc=$2; fam=$3; nd=$4; shift 4
  nvcc -I$inc -DBENCH_FAMILY=$fam -DBENCH_NDIM=$nd "$@" $common -o _build/r2_$name &
  n=$((n+1)); if [ $((n % 8)) = 0 ]; then wait; fi
}
for mb in 3 4 5 6; do
  b old_osc8_mb$mb _build/old_include Oscillatory 8 -DPGN_EVAL_MIN_BLOCKS=$mb
  b new_osc8_mb$mb $R/include Oscillatory 8 -DPGN_EVAL_MIN_BLOCKS=$mb "$@"
done
b new_osc8_w8mb2 $R/include Oscillatory 8 -DPGN_EVAL_WARPS=8 -DPGN_EVAL_MIN_BLOCKS=2 "$@"
b new_osc8_w2mb8 $R/include Oscillatory 8 -DPGN_EVAL_WARPS=2 -DPGN_EVAL_MIN_BLOCKS=8 "$@"
for fam in "Gaussian 8 3" "ProductPeak 6 4" "CZero 10 4" "Discontinuous 10 4" "Gaussian 5 3" "CornerPeak 8 8"; do
  set -- $fam
  b old_$1$2 _build/old_include $1 $2 -DPGN_EVAL_MIN_BLOCKS=$3
  b new_$1$2 $R/include $1 $2 -DPGN_EVAL_MIN_BLOCKS=$3
done
wait
ls _build | grep -c r2_
