#!/bin/bash
# round 2, session H: evaluation-kernel variants of the working tree (cross-compiled here, run by tools/gpu_sessions/r2_h.sh)
#   seg*    PGN_EVAL_SEGMENTS=1 (compile-time trip runs + chain-A-only last trip), noseg* the per-trip set lookup
#   *_np    PGN_EVAL_PREFETCH_BOUNDS=0;  old_* the round-1 kernel (include/ tree of commit 55affdb)
cd "$(dirname "$0")"
mkdir -p _build/old_include
rm -f _build/r2h_*
R=../..
(cd $R && git archive 55affdb include) | tar -x -C _build/old_include --strip-components=1
common="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DPGN_EVAL_NO_GENZ_TUNING -w eval_bench.cu -L$R/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../gpuintegration_b200"
n=0
b() { # name include-dir family ndim flags...
  name=$1; inc=$2; fam=$3; nd=$4; shift 4
  nvcc -I$inc -DBENCH_FAMILY=$fam -DBENCH_NDIM=$nd "$@" $common -o _build/r2h_$name &
  n=$((n+1)); if [ $((n % 8)) = 0 ]; then wait; fi
}
for mb in 3 4 5 6; do
  b seg_osc8_mb$mb $R/include Oscillatory 8 -DPGN_EVAL_MIN_BLOCKS=$mb
done
b noseg_osc8_mb4 $R/include Oscillatory 8 -DPGN_EVAL_MIN_BLOCKS=4 -DPGN_EVAL_SEGMENTS=0
b noseg_osc8_mb5 $R/include Oscillatory 8 -DPGN_EVAL_MIN_BLOCKS=5 -DPGN_EVAL_SEGMENTS=0
b seg_osc8_mb4_np $R/include Oscillatory 8 -DPGN_EVAL_MIN_BLOCKS=4 -DPGN_EVAL_PREFETCH_BOUNDS=0
for fam in "Gaussian 8 3" "ProductPeak 6 4" "CZero 10 4" "Discontinuous 10 4" "Gaussian 5 3" "CornerPeak 8 8" "Oscillatory 5 4" "Oscillatory 6 4"; do
  set -- $fam
  b old_$1$2 _build/old_include $1 $2 -DPGN_EVAL_MIN_BLOCKS=$3
  b seg_$1$2 $R/include $1 $2 -DPGN_EVAL_MIN_BLOCKS=$3
  b seg_$1$2_np $R/include $1 $2 -DPGN_EVAL_MIN_BLOCKS=$3 -DPGN_EVAL_PREFETCH_BOUNDS=0
  b noseg_$1$2_np $R/include $1 $2 -DPGN_EVAL_MIN_BLOCKS=$3 -DPGN_EVAL_PREFETCH_BOUNDS=0 -DPGN_EVAL_SEGMENTS=0
done
b seg_Gaussian8_mb4 $R/include Gaussian 8 -DPGN_EVAL_MIN_BLOCKS=4
b seg_CornerPeak8_mb6 $R/include CornerPeak 8 -DPGN_EVAL_MIN_BLOCKS=6
b seg_CornerPeak8_mb5_np $R/include CornerPeak 8 -DPGN_EVAL_MIN_BLOCKS=5 -DPGN_EVAL_PREFETCH_BOUNDS=0
wait
ls _build | grep -c r2h_
