#!/bin/bash
# round 2, session AE (1 GPU): the default bench command (every side leg) on the final library
mkdir -p gpurun_out
( time timeout 235 python bench.py ) > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2ae_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ae_bench.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g roof %.3f parity %s'%(d['value'],d['e2e']['value'],d['roofline']['frac'], d['parity_ok']), d['stage_ms_per_step'])
for c in d.get('configs', []):
    print(c['config'], 'ours %.0f ms' % c['ours']['ms'], 'status', c['ours']['status'], 'ref gpu %.0f ms' % c['reference_gpu']['ms'] if c.get('reference_gpu') else '', 'speedup', c.get('speedup_vs_reference_gpu'), 'parity', c.get('parity', {}).get('ok'))
print('ladder', json.dumps(d.get('native_ladder'))[:800])
print('ref gpu speedup', d.get('reference_gpu', {}).get('speedup_e2e'), 'cpu', d.get('cpu_baseline'))
PY
