#!/bin/bash
# TEST INFRASTRUCTURE.  Compiles the REFERENCE'S OWN unit tests, unmodified and from where they lie under
# $REFERENCE, against this project's headers: the only change with respect to the reference's build is the include
# path (include/pagani_b200/compat and include/ come first, so "cuda/pagani/quad/GPUquad/*.cuh" resolve to the
# shims on the B200-native engine) and the library they link.  Outputs go to tests/cpp/_build/ref_tests/ (git-ignored,
# travels to the GPU box); tests/test_gpu_reference_tests.py runs them.  No reference source is copied.
set -e
cd "$(dirname "$0")"
REFERENCE=${REFERENCE:-/root/reference}
ROOT=../..
OUT=_build/ref_tests
# the same programs built against the REFERENCE'S OWN headers: the yardstick the swapped builds are compared with
# (a reference test that fails on its own arithmetic -- Uniform_sub_division.cu:64 sums 243 x (1/3)^5 and expects
# exactly 1 -- must fail the same way, no more and no less)
REFOUT=_build/ref_tests_reference
if [ ! -d "$REFERENCE/test/cuda/pagani" ]; then
  echo "reference sources not present: keeping prebuilt $OUT (if any)"; exit 0
fi
mkdir -p $OUT $REFOUT
REFFLAGS="-std=c++17 -O2 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -w -I$REFERENCE \
  -I$REFERENCE/externals -I/usr/local/cuda/include/nvtx3"
FLAGS="-std=c++17 -O2 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -w \
  -I$ROOT/include/pagani_b200/compat -I$ROOT/include -I$REFERENCE -I$REFERENCE/externals \
  -L$ROOT/gpuintegration_b200 -lpagani_b200 -Xlinker -rpath -Xlinker \$ORIGIN/../../../../gpuintegration_b200"
T=$REFERENCE/test/cuda/pagani
newer() { # target deps...: true when the target is missing or older than a dependency
  local t=$1; shift
  [ ! -e "$t" ] && return 0
  for d in "$@"; do [ "$d" -nt "$t" ] && return 0; done
  return 1
}
DEPS="$ROOT/gpuintegration_b200/libpagani_b200.so $ROOT/include/pagani_b200/stages.cuh $ROOT/include/pagani_b200/Workspace.cuh $ROOT/include/pagani_b200/evaluate.cuh $ROOT/include/pagani_b200.h"
pids=()
for t in RegionSplitting RegionFiltering RegionSampling Uniform_sub_division Easy_Integrals BoxIntegrals NanIntegral \
         pagani_loop_test accuracy_improves_with_epsrel; do
  if newer $OUT/$t $DEPS $T/$t.cu; then
    nvcc $FLAGS $T/$t.cu -o $OUT/$t &
    pids+=($!)
  fi
  if newer $REFOUT/$t $T/$t.cu; then
    nvcc $REFFLAGS $T/$t.cu -o $REFOUT/$t &
    pids+=($!)
  fi
done
# the C linkage test: a C main calling the reference's extern "C" wrapper around Workspace::integrate
if newer $OUT/test_integration_from_c $DEPS $T/do_integration_from_c.cu $T/test_integration_from_c.c; then
  ( nvcc $FLAGS -c $T/do_integration_from_c.cu -o $OUT/do_integration_from_c.o &&
    gcc -c -I$REFERENCE $T/test_integration_from_c.c -o $OUT/test_integration_from_c.o &&
    nvcc $FLAGS $OUT/do_integration_from_c.o $OUT/test_integration_from_c.o -o $OUT/test_integration_from_c ) &
  pids+=($!)
fi
# this project's own test of stateful functors (on the reference's cudaArray / Interp1D helpers) and snapshot / resume
if newer $OUT/test_stateful_resume $DEPS test_stateful_resume.cu; then
  nvcc $FLAGS test_stateful_resume.cu -o $OUT/test_stateful_resume &
  pids+=($!)
fi
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
ls $OUT | grep -v "\.o$" | tr '\n' ' '; echo
exit $rc
