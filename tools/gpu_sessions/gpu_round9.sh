#!/bin/bash
# session 8, call 2: second cos sweep (four-point trips, launch bounds), full GPU tests, bench
mkdir -p gpurun_out
: > gpurun_out/cos_sweep2.jsonl
for b in tools/evalbench/_build/cos_*; do
  for parts in 5; do
    echo -n "{\"bin\": \"$(basename $b)\", \"run\": " >> gpurun_out/cos_sweep2.jsonl
    timeout 60 $b $parts 7 | tr -d '\n' >> gpurun_out/cos_sweep2.jsonl; echo "}" >> gpurun_out/cos_sweep2.jsonl
  done
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/cos_sweep2.jsonl') if l.strip().endswith('}}')]
for r in sorted(rows,key=lambda r:(r['run']['regions'],-r['run']['evals_per_s'])):
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g roof %.3f'%(d['value'],d['e2e']['value'],d['roofline']['frac']), d['stage_ms_per_step'], d.get('roofline_hbm'))
PY
