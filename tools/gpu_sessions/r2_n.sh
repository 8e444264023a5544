#!/bin/bash
# round 2, session N (1 GPU): compaction capped at 64 registers (four resident blocks) -- stage bench, parity + stage
# tests, short bench
mkdir -p gpurun_out
timeout 120 python tools/stage_bench.py shipped > gpurun_out/r2n_stages.jsonl 2> gpurun_out/r2n_stages.err; echo "stage bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2n_stages.jsonl'):
    r=json.loads(l)
    if r['regions'] >= 2_000_000:
        print(r['label'], r['epsrel'], r['regions'], r['survivors'], 'classify %.3f ms %.2f' % (r['classify_ms'], r['classify_frac']), 'split %.3f ms %.2f' % (r['split_ms'], r['compact_split_frac'] or 0), 'heur %.3f' % r['heuristic_ms'])
PY
timeout 900 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_stages.py tests/test_gpu_integrate.py -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_pytest.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-side-legs > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2n_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
PY
