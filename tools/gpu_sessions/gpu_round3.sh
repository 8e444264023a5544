#!/bin/bash
mkdir -p gpurun_out
B=tools/evalbench/_build
( for exe in eb_Oscillatory8_old eb_Oscillatory8_mb1 eb_Oscillatory8_mb8 eb_CZero8_old eb_CZero8_mb1 eb_CZero8_mb8 \
             eb_Gaussian8_old eb_Gaussian8_mb1 eb_Gaussian8_mb8 eb_ProductPeak6_mb1 eb_ProductPeak6_mb8; do
    timeout 120 $B/$exe 5 4; done
  for exe in eb_CZero10_old eb_CZero10_mb1 eb_CZero10_mb8; do timeout 120 $B/$exe 3 4; done
  timeout 120 $B/eb_ProductPeak6_mb1 8 4; timeout 120 $B/eb_ProductPeak6_mb8 8 4
) > gpurun_out/evalbench.jsonl 2>&1
cat gpurun_out/evalbench.jsonl
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_multirank.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q > gpurun_out/pytest_multirank.log 2>&1
echo "pytest multirank rc=$?" >> gpurun_out/pytest_multirank.log
tail -5 gpurun_out/pytest_multirank.log
timeout 600 python bench.py --no-side-legs > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','stage_ms_per_step')}, d['roofline']['frac'], d['e2e']['value'])
PY
