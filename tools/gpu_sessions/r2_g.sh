#!/bin/bash
# round 2, session G (1 GPU): classify fast path + compact 8 axes (stage bench), evaluation kernel with the cheaper
# prologue, full GPU suite, bench with zero-copy statistics + host timing, ncu of the shipped evaluation kernel and stages
mkdir -p gpurun_out
timeout 120 python tools/stage_bench.py shipped > gpurun_out/r2g_stages.jsonl 2> gpurun_out/r2g_stages.err; echo "stage bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2g_stages.jsonl'):
    r=json.loads(l)
    if r['regions'] >= 4_000_000:
        print(r['label'], r['epsrel'], r['regions'], r['survivors'], 'classify %.3f ms %.2f' % (r['classify_ms'], r['classify_frac']), 'split %.3f ms %.2f' % (r['split_ms'], r['compact_split_frac'] or 0), 'heur %.3f' % r['heuristic_ms'])
PY
: > gpurun_out/r2g_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2g_evalbench.jsonl; timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2g_evalbench.jsonl; echo "}" >> gpurun_out/r2g_evalbench.jsonl; }
for b in r2_new_osc8_mb4 r2_new_osc8_mb5 r2_new_osc8_mb6 r2_old_osc8_mb4; do run $b 5; done
for b in r2_new_CZero10 r2_new_Discontinuous10 r2_old_Discontinuous10; do run $b 3; done
for b in r2_new_Gaussian8 r2_old_Gaussian8 r2_new_CornerPeak8 r2_old_CornerPeak8; do run $b 4; done
for b in r2_new_ProductPeak6 r2_old_ProductPeak6; do run $b 8; done
for b in r2_new_Gaussian5 r2_old_Gaussian5; do run $b 12; done
python - <<'PY'
import json
for l in open('gpurun_out/r2g_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
timeout 1500 python -m pytest -q -m gpu tests > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2g_pytest.log
PGN_HOST_TIMING=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-side-legs > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
tail -6 gpurun_out/r2g_bench.err
PGN_NO_ZERO_COPY_STATS=1 timeout 300 python bench.py --steps 3 --warmup 2 --no-side-legs > gpurun_out/r2g_bench_nozc.json 2> /dev/null
python - <<'PY'
import json
for f in ('r2g_bench', 'r2g_bench_nozc'):
    d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
    print(f, 'value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, 'gap %.2f' % (d['ms_per_step']-sum(d['stage_ms_per_step'].values())))
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2g_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2g_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_classify|k_compact' -o gpurun_out/r2g_stages_1e-4 python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2g_ncu_stages.log 2>&1; echo "ncu stages rc=$?"
