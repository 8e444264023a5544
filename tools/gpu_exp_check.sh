#!/bin/bash
# exp_pinned: equality with exp() (strict), evaluation-kernel timing with it, and the exp-family parity tests against
# a library built with PGN_GENZ_FAST_EXP=1 (swapped in on the GPU box only).  Build first: tools/evalbench/build_exp.sh
mkdir -p gpurun_out
timeout 120 tests/cpp/_build/test_fastmath 20000000 pe > gpurun_out/fastmath_exp.txt 2>&1; echo "fastmath rc=$?"; grep -E "pinned|cos variant" gpurun_out/fastmath_exp.txt
: > gpurun_out/exp_sweep.jsonl
for b in tools/evalbench/_build/exp_*; do
  [ -f "$b" ] || continue
  case $b in *10_*) parts=2;; *) parts=5;; esac
  echo -n "{\"bin\": \"$(basename $b)\", \"run\": " >> gpurun_out/exp_sweep.jsonl; timeout 60 $b $parts 5 | tr -d '\n' >> gpurun_out/exp_sweep.jsonl; echo "}" >> gpurun_out/exp_sweep.jsonl
done
python - <<'PY'
import json
for l in open('gpurun_out/exp_sweep.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:80]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
cp tools/evalbench/_build/fastexp/libpagani_b200.so gpuintegration_b200/libpagani_b200.so
timeout 900 python -m pytest -q -x -m gpu tests/test_gpu_parity.py tests/test_gpu_search.py -k "f4 or f5 or f6 or 4-8 or 5-10 or 6-8 or 5-5 or 6-10 or 6-7 or 4-7 or 5-8 or 6-6" > gpurun_out/pytest_fastexp.log 2>&1; echo "pytest fastexp rc=$?"; tail -5 gpurun_out/pytest_fastexp.log
