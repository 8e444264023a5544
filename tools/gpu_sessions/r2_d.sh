#!/bin/bash
# round 2, session D: fourth-difference fix (xy_2 golden), fallback tests, goldens for the 3-D / 4-D multilinear cases,
# evaluation-kernel timings of the final functor set, full bench with side legs (how long does it take?)
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu tests/test_gpu_reference_tests.py tests/test_gpu_dropin.py > gpurun_out/r2d_pytest_reftests.log 2>&1; echo "reference tests rc=$?"; tail -25 gpurun_out/r2d_pytest_reftests.log
timeout 600 python -m pytest -q -m gpu tests/test_gpu_parity.py > gpurun_out/r2d_pytest_parity.log 2>&1; echo "parity rc=$?"; tail -6 gpurun_out/r2d_pytest_parity.log
rm -rf gpurun_out/golden2
timeout 300 python - <<'PY' > gpurun_out/r2d_golden.log 2>&1
import sys
sys.path.insert(0, "tools")
import make_golden_trajectories as m
m.CASES = [c for c in m.CASES if c[0] == 7 and c[1] in (3, 4)]
sys.argv = ["x", "--out", "gpurun_out/golden2"]
m.main()
PY
echo "golden rc=$?"; tail -3 gpurun_out/r2d_golden.log; rm -rf gpurun_out/golden2/scratch
: > gpurun_out/r2d_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2d_evalbench.jsonl; timeout 60 tools/evalbench/_build/$1 $2 3 | tr -d '\n' >> gpurun_out/r2d_evalbench.jsonl; echo "}" >> gpurun_out/r2d_evalbench.jsonl; }
for b in r2_new_osc8_mb4 r2_old_osc8_mb4; do run $b 5; done
for b in r2_new_Gaussian8 r2_old_Gaussian8 r2_new_CornerPeak8 r2_old_CornerPeak8; do run $b 4; done
for b in r2_new_ProductPeak6 r2_old_ProductPeak6; do run $b 8; done
for b in r2_new_CZero10 r2_old_CZero10 r2_new_Discontinuous10 r2_old_Discontinuous10; do run $b 3; done
for b in r2_new_Gaussian5 r2_old_Gaussian5; do run $b 12; done
python - <<'PY'
import json
for l in open('gpurun_out/r2d_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
( time timeout 900 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2d_bench.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity_ok', d['parity_ok'])
print('roofline', {k:d['roofline'][k] for k in ('frac','frac_executed','fp64_pipe_pct','achieved','peak')})
for c in d.get('configs', []):
    print(c['config'], 'ours ms', round(c['ours']['ms'],1), 'status', c['ours']['status'], 'parity', c['parity']['ok'], c['parity'].get('mismatches'), 'ref ms', c.get('reference_gpu',{}).get('ms'), 'speedup', c.get('speedup_vs_reference_gpu'), 'cpu', c.get('cpu_cuhre_same_epsrel',{}).get('seconds'), c.get('cpu_cuhre_same_epsrel',{}).get('status'))
l=d.get('native_ladder')
if l:
    print('ladder', l['stopped_by'], l['tightest_epsrel_reached'])
    for s in l['steps']: print('  ', s['epsrel'], s['status'], round(s['ms'],1), s['max_regions_in_flight'], '%.3g'%s['evals_per_s'], '%.2e'%s['rel_err_vs_true'])
print('ref gpu', d.get('reference_gpu',{}).get('speedup_e2e'), [ (x['epsrel'], x['ms']) for x in d.get('reference_gpu',{}).get('per_epsrel',[])])
print('hbm', json.dumps(d.get('roofline_hbm'))[:600])
PY
