#!/bin/bash
# round 2, session M (1 GPU): the shipped state -- full GPU suite, smoke, stage bench, bench with every leg, the
# reference arm, ncu launch list of the bench, ncu --set full of the evaluation kernel and of classify / compact
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2m_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"
timeout 120 python tools/stage_bench.py shipped > gpurun_out/r2m_stages.jsonl 2> gpurun_out/r2m_stages.err; echo "stage bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2m_stages.jsonl'):
    r=json.loads(l)
    if r['regions'] >= 2_000_000:
        print(r['label'], r['epsrel'], r['regions'], r['survivors'], 'classify %.3f ms %.2f' % (r['classify_ms'], r['classify_frac']), 'split %.3f ms %.2f' % (r['split_ms'], r['compact_split_frac'] or 0), 'heur %.3f' % r['heuristic_ms'])
PY
( time timeout 1200 python bench.py ) > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2m_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2m_bench.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g roof %.3f parity %s'%(d['value'],d['e2e']['value'],d['roofline']['frac'], d['parity_ok']), d['stage_ms_per_step'])
for c in d.get('configs', []):
    print(c['config'], 'ours %.0f ms' % c['ours']['ms'], 'status', c['ours']['status'], 'ref gpu %.0f ms' % c['reference_gpu']['ms'] if c.get('reference_gpu') else '', 'speedup', c.get('speedup_vs_reference_gpu'), 'parity', c.get('parity', {}).get('ok'))
print('ladder', json.dumps(d.get('native_ladder'))[:1500])
print('ref gpu speedup', d.get('reference_gpu', {}).get('speedup_e2e'))
PY
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/r2m_bench_ref.json 2> gpurun_out/r2m_bench_ref.err; echo "bench ref rc=$?"; cut -c1-200 gpurun_out/r2m_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2m_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-side-legs > gpurun_out/r2m_bench_ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/r2m_launches_bench.csv > gpurun_out/r2m_launches_summary.txt 2>&1; cat gpurun_out/r2m_launches_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2m_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'k_classify|k_compact' --launch-skip 26 -c 8 -o gpurun_out/r2m_stages_1e-4 python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2m_ncu_stages.log 2>&1; echo "ncu stages rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_search -s 6 -c 2 -o gpurun_out/r2m_search python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2m_ncu_search.log 2>&1; echo "ncu search rc=$?"
ls -la gpurun_out | grep r2m
