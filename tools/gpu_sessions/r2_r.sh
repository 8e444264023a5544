#!/bin/bash
# round 2, session R (1 GPU): the final library and the rebuilt reference test programs -- full GPU suite, smoke,
# bench without side legs
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2r_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2r_smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2r_bench.json') if l.startswith('{')][-1])
print('bench value %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'parity', d['parity_ok'], {k: round(v,2) for k,v in d['stage_ms_per_step'].items()}, 'fp64 pipe', d['roofline'].get('fp64_pipe_pct'), 'frac', round(d['roofline']['frac'],3))
PY
