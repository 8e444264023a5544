#!/bin/bash
# round 2, session AB (1 GPU): the library with the warp-uniform warp index, the range-promise kernels (F_1, F_5) and the
# cosine's parity predicate -- evalbench of the shipped tuning, full GPU suite, smoke, bench, ncu of the evaluation kernel
mkdir -p gpurun_out
: > gpurun_out/r2ab_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2ab_evalbench.jsonl; timeout 120 tools/evalbench/_build/r2ab_$1 $2 3 | tr -d '\n' >> gpurun_out/r2ab_evalbench.jsonl; echo "}" >> gpurun_out/r2ab_evalbench.jsonl; }
for rep in 1 2; do run c010_chk 3; run c010_nochk 3; run osc8_ship 5; run osc5_ship 12; done
python - <<'PY'
import json
for l in open('gpurun_out/r2ab_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2ab_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2ab_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ab_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, 'fp64 pipe', d['roofline'].get('fp64_pipe_pct'), 'frac', round(d['roofline']['frac'],3))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2ab_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2ab_ncu.log 2>&1; echo "ncu rc=$?"
