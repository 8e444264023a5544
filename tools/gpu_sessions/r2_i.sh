#!/bin/bash
# round 2, session I (1 GPU): sibling-pair classify with TMA-staged rows (k_classify_pairs) against the one-region
# kernel (PGN_CLASSIFY_NO_PAIRS=1): full GPU suite on the pairs kernel, stage bench both ways, bench, ncu of both
mkdir -p gpurun_out
timeout 1500 python -m pytest -q -m gpu tests -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2i_pytest.log
for v in pairs nopairs; do
  if [ $v = nopairs ]; then export PGN_CLASSIFY_NO_PAIRS=1; else unset PGN_CLASSIFY_NO_PAIRS; fi
  timeout 120 python tools/stage_bench.py $v > gpurun_out/r2i_stages_$v.jsonl 2> gpurun_out/r2i_stages_$v.err; echo "stage bench $v rc=$?"
done
unset PGN_CLASSIFY_NO_PAIRS
python - <<'PY'
import json
for v in ('pairs', 'nopairs'):
    for l in open('gpurun_out/r2i_stages_%s.jsonl' % v):
        r=json.loads(l)
        if r['regions'] >= 2_000_000:
            print(r['label'], r['epsrel'], r['regions'], r['survivors'], 'classify %.3f ms %.2f' % (r['classify_ms'], r['classify_frac']), 'split %.3f ms %.2f' % (r['split_ms'], r['compact_split_frac'] or 0), 'heur %.3f' % r['heuristic_ms'])
PY
timeout 300 python bench.py --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2i_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2i_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'k_classify' --launch-skip 13 -c 4 -o gpurun_out/r2i_classify_pairs python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2i_ncu_stages.log 2>&1; echo "ncu stages rc=$?"
ls -la gpurun_out
