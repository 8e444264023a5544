#!/bin/bash
# round 2, session F (1 GPU): reference test programs, evaluation-kernel variants (bounds prefetch, launch bounds),
# HBM stages: pipelined classify (depth x launch bound), compact with 4 / 8 axes in flight, ncu of the shipped stages
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_reference_tests.py tests/test_gpu_stages.py tests/test_gpu_dropin.py > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2f_pytest.log
: > gpurun_out/r2f_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2f_evalbench.jsonl; timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2f_evalbench.jsonl; echo "}" >> gpurun_out/r2f_evalbench.jsonl; }
for b in r2_new_osc8_mb3 r2_new_osc8_mb4 r2_new_osc8_mb5 r2_new_osc8_mb6 r2_new_osc8_w8mb2 r2_old_osc8_mb4; do run $b 5; done
for b in r2_new_CZero10 r2_old_CZero10 r2_new_Discontinuous10; do run $b 3; done
for b in r2_new_Gaussian8 r2_old_Gaussian8; do run $b 4; done
python - <<'PY'
import json
for l in open('gpurun_out/r2f_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
: > gpurun_out/r2f_stages.jsonl
cp gpuintegration_b200/libpagani_b200.so /tmp/shipped.so
timeout 120 python tools/stage_bench.py shipped >> gpurun_out/r2f_stages.jsonl
for v in cls_d2_b2 cls_d2_b3 cls_d3_b2 cls_d3_b3 cls_d4_b2 cmp_a4 cmp_a8; do
  cp tools/evalbench/_build/$v/libpagani_b200.so gpuintegration_b200/libpagani_b200.so
  timeout 120 python tools/stage_bench.py $v >> gpurun_out/r2f_stages.jsonl
done
cp /tmp/shipped.so gpuintegration_b200/libpagani_b200.so
python - <<'PY'
import json
for l in open('gpurun_out/r2f_stages.jsonl'):
    r=json.loads(l)
    if r['regions'] >= 8_000_000:
        print(r['label'], r['epsrel'], r['regions'], r['survivors'], 'classify %.3f ms %.2f' % (r['classify_ms'], r['classify_frac']), 'split %.3f ms %.2f' % (r['split_ms'], r['compact_split_frac'] or 0))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_compact' -o gpurun_out/r2f_stages_1e-4 python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2f_ncu_stages.log 2>&1; echo "ncu stages rc=$?"; tail -6 gpurun_out/r2f_ncu_stages.log
