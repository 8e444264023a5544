#!/bin/bash
# round 2, session AD (1 GPU): the final library of the round -- full GPU suite, smoke, bench without side legs, ncu launch
# list of the bench command
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2ad_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ad_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2ad_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ad_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, 'fp64 pipe', d['roofline'].get('fp64_pipe_pct'), 'frac', round(d['roofline']['frac'],3), 'frac_executed', round(d['roofline']['frac_executed'],3))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2ad_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-side-legs > gpurun_out/r2ad_bench_ncu.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/r2ad_launches_bench.csv > gpurun_out/r2ad_launches_summary.txt 2>&1; head -30 gpurun_out/r2ad_launches_summary.txt
