#!/bin/bash
# round 2, session T (1 GPU): oscillatory family with per-lane weights from shared memory -- parity + integrate +
# drop-in tests, bench, ncu --set full of the shipped evaluation kernel (for profiles/ncu_eval_f1_8d.json)
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_integrate.py tests/test_gpu_dropin.py tests/test_gpu_stages.py -x > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2t_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, 'frac', round(d['roofline']['frac'],3))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2t_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc=$?"
