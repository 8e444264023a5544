#!/bin/bash
# session 8, call: full GPU tests, smoke, bench (all legs), reference arm, BASELINE configs, ncu launch list of the bench
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g roof %.3f'%(d['value'],d['e2e']['value'],d['roofline']['frac']), d['stage_ms_per_step'], json.dumps(d.get('roofline_hbm')))
PY
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-side-legs > gpurun_out/bench_ncu.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_bench.csv > gpurun_out/launches_summary.txt 2>&1; cat gpurun_out/launches_summary.txt
