#!/bin/bash
# session 8, call 1: cos_pinned equality + variant sweep, full GPU tests, bench, ncu of evaluate / classify / compact
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
for v in 1 2 3 4; do timeout 120 tests/cpp/_build/test_fastmath_v$v 20000000 > gpurun_out/fastmath_v$v.txt 2>&1; echo "fastmath v$v rc=$?"; tail -1 gpurun_out/fastmath_v$v.txt; done
: > gpurun_out/cos_sweep.jsonl
for b in tools/evalbench/_build/cos_*; do
  for parts in 5 6; do
    echo -n "{\"bin\": \"$(basename $b)\", \"run\": " >> gpurun_out/cos_sweep.jsonl
    timeout 60 $b $parts 5 | tr -d '\n' >> gpurun_out/cos_sweep.jsonl; echo "}" >> gpurun_out/cos_sweep.jsonl
  done
done
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/cos_sweep.jsonl') if l.strip().endswith('}}')]
for r in sorted(rows,key=lambda r:(r['run']['regions'],-r['run']['evals_per_s'])):
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g roof %.3f'%(d['value'],d['e2e']['value'],d['roofline']['frac']), d['stage_ms_per_step'], d.get('roofline_hbm'))
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:evaluate_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/eval_f1_8d_s8 python tools/ncu_eval.py 1 8 4 3 > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'k_classify|k_compact' --launch-skip 24 --launch-count 7 -f -o gpurun_out/stages_s8 python tools/ncu_stages.py > gpurun_out/ncu_stages.log 2>&1; echo "ncu stages rc=$?"; tail -6 gpurun_out/ncu_stages.log
