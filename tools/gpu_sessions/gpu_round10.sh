#!/bin/bash
# session 8, call 4: FP64 issue/latency micro-benchmarks, full GPU tests, bench, BASELINE configs side by side, ncu of the default evaluation kernel
mkdir -p gpurun_out
timeout 120 tools/evalbench/_build/fp64_lat > gpurun_out/fp64_lat.jsonl 2>&1; echo "fp64_lat rc=$?"; cat gpurun_out/fp64_lat.jsonl
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
CPU_CAP_S=5 timeout 1200 python tools/baseline_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; tail -2 gpurun_out/configs.err
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    d=json.loads(l)
    if 'ours' not in d: print(d); continue
    o=d['ours']; r=d.get('reference_gpu',{}); c=d.get('cpu_cuhre_restated',{})
    print(d['config'][:28], d['epsrel'], d['mem_mode'], 'ours %.1f ms st %d nreg %d %.3g ev/s roof %.2f'%(o['ms'],o['status'],o['nregions'],o['evals_per_s'],o['roofline_frac_algorithmic'] or 0),
          '| ref %s ms same %s x%.2f'%(r.get('ms'),d.get('same_nregions_status'),d.get('speedup_vs_reference_gpu',0)), '| cpu %.3g ev/s'%c.get('evals_per_s',0))
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:evaluate_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/eval_f1_8d_s8b python tools/ncu_eval.py 1 8 4 3 > gpurun_out/ncu_eval.log 2>&1; echo "ncu eval rc=$?"
