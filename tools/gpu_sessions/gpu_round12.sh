#!/bin/bash
# last call of the round (few GPU-minutes left): pow2_pinned equality, evaluation-kernel timing with it, and the
# pow-family parity tests against a library built with PGN_GENZ_FAST_POW=1 (swapped in on the box only)
mkdir -p gpurun_out
timeout 60 tests/cpp/_build/test_fastmath 4000000 pow > gpurun_out/fastmath_pow.txt 2>&1; echo "fastmath rc=$?"; grep -E "pow|cos variant" gpurun_out/fastmath_pow.txt
: > gpurun_out/pow_sweep.jsonl
for b in tools/evalbench/_build/pow_Gaussian8_* ; do echo -n "{\"bin\": \"$(basename $b)\", \"run\": " >> gpurun_out/pow_sweep.jsonl; timeout 30 $b 4 3 | tr -d '\n' >> gpurun_out/pow_sweep.jsonl; echo "}" >> gpurun_out/pow_sweep.jsonl; done
for b in tools/evalbench/_build/pow_ProductPeak6_* ; do echo -n "{\"bin\": \"$(basename $b)\", \"run\": " >> gpurun_out/pow_sweep.jsonl; timeout 30 $b 8 3 | tr -d '\n' >> gpurun_out/pow_sweep.jsonl; echo "}" >> gpurun_out/pow_sweep.jsonl; done
python - <<'PY'
import json
for l in open('gpurun_out/pow_sweep.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:80]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
cp tools/evalbench/_build/fastpow/libpagani_b200.so gpuintegration_b200/libpagani_b200.so
timeout 170 python -m pytest -q -x -m gpu "tests/test_gpu_parity.py::test_trajectory_matches_golden[f4_5_0.001]" "tests/test_gpu_parity.py::test_trajectory_matches_golden[f2_6_0.001]" "tests/test_gpu_parity.py::test_trajectory_matches_live_reference[4-8-1e-05]" "tests/test_gpu_parity.py::test_trajectory_matches_live_reference[2-6-1e-06]" "tests/test_gpu_parity.py::test_point_loop_variants_agree[4-8]" "tests/test_gpu_parity.py::test_point_loop_variants_agree[2-6]" "tests/test_gpu_parity.py::test_trajectory_matches_live_reference[4-8-1e-09]" > gpurun_out/pytest_fastpow.log 2>&1; echo "pytest fastpow rc=$?"; tail -5 gpurun_out/pytest_fastpow.log
