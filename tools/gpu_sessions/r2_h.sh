#!/bin/bash
# round 2, session H (1 GPU): evaluation kernel with compile-time trip segments + chain-A-only last trip, bounds
# prefetch on/off per family, restored classify (depth 3); parity suite on the new library; ncu of the shipped kernels
# (captures bounded so that gpurun_out stays below the 64 MiB that travel back)
mkdir -p gpurun_out
: > gpurun_out/r2h_evalbench.jsonl
run() { echo -n "{\"bin\": \"$1\", \"run\": " >> gpurun_out/r2h_evalbench.jsonl; timeout 60 tools/evalbench/_build/r2h_$1 $2 3 | tr -d '\n' >> gpurun_out/r2h_evalbench.jsonl; echo "}" >> gpurun_out/r2h_evalbench.jsonl; }
for b in seg_osc8_mb3 seg_osc8_mb4 seg_osc8_mb5 seg_osc8_mb6 noseg_osc8_mb4 noseg_osc8_mb5 seg_osc8_mb4_np; do run $b 5; done
for f in CZero10 Discontinuous10; do for v in old seg seg_np noseg_np; do run ${v%_np}_$f${v#${v%_np}} 3; done; done
for f in Gaussian8 CornerPeak8; do for v in old seg seg_np noseg_np; do run ${v%_np}_$f${v#${v%_np}} 4; done; done
run seg_Gaussian8_mb4 4; run seg_CornerPeak8_mb6 4; run seg_CornerPeak8_mb5_np 4
for f in ProductPeak6 Oscillatory6; do for v in old seg seg_np noseg_np; do run ${v%_np}_$f${v#${v%_np}} 8; done; done
for f in Gaussian5 Oscillatory5; do for v in old seg seg_np noseg_np; do run ${v%_np}_$f${v#${v%_np}} 12; done; done
python - <<'PY'
import json
for l in open('gpurun_out/r2h_evalbench.jsonl'):
    try: r=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(r['bin'], r['run']['regions'], '%.4g'%r['run']['evals_per_s'], '%.17g'%r['run']['estimate'], '%.17g'%r['run']['errorest'])
PY
timeout 900 python -m pytest -q -m gpu tests/test_gpu_parity.py tests/test_gpu_integrate.py tests/test_gpu_stages.py tests/test_gpu_dropin.py > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2h_pytest.log
timeout 120 python tools/stage_bench.py shipped > gpurun_out/r2h_stages.jsonl 2> gpurun_out/r2h_stages.err; echo "stage bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/r2h_stages.jsonl'):
    r=json.loads(l)
    if r['regions'] >= 4_000_000:
        print(r['label'], r['epsrel'], r['regions'], r['survivors'], 'classify %.3f ms %.2f' % (r['classify_ms'], r['classify_frac']), 'split %.3f ms %.2f' % (r['split_ms'], r['compact_split_frac'] or 0), 'heur %.3f' % r['heuristic_ms'])
PY
timeout 300 python bench.py --steps 3 --warmup 2 --no-side-legs > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2h_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()})
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel_bp -s 1 -c 1 -o gpurun_out/r2h_eval_f1_8d python tools/ncu_eval.py 1 8 4 2 > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'k_classify|k_compact' --launch-skip 22 -c 10 -o gpurun_out/r2h_stages_1e-4 python tools/ncu_stages.py 1 8 1e-4 > gpurun_out/r2h_ncu_stages.log 2>&1; echo "ncu stages rc=$?"
ls -la gpurun_out
