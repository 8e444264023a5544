#!/bin/bash
# round 2, session O (8 GPUs): the bench exactly as the driver launches it at N=8, every side leg included
# (configs 3-5 partitioned over 8 ranks, native-memory tolerance ladder on 8 x 180 GB)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err; echo "bench n=8 rc=$?"
tail -4 gpurun_out/r2o_bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2o_bench_n8.json') if l.startswith('{')][-1])
print('value %.4g ms/step %.1f e2e %.4g parity %s'%(d['value'],d['ms_per_step'],d['e2e']['value'], d['parity_ok']), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()})
for c in d.get('configs', []):
    print(c['config'], 'ours %.0f ms' % c['ours']['ms'], 'status', c['ours']['status'], 'ref gpu %.0f ms' % c['reference_gpu']['ms'] if c.get('reference_gpu') and 'ms' in c['reference_gpu'] else '', 'speedup', c.get('speedup_vs_reference_gpu'), 'parity', c.get('parity', {}).get('ok'))
l=d.get('native_ladder') or {}
print('ladder reached', l.get('tightest_epsrel_reached'), 'stopped', l.get('stopped_by'))
for s in l.get('steps', []):
    print('  eps %.3g status %d ms %.1f nregions %d in flight %d relerr %.2e' % (s['epsrel'], s['status'], s['ms'], s['nregions'], s['max_regions_in_flight'], s['rel_err_vs_true']))
print('cpu_baseline cores', d['cpu_baseline']['cores'])
PY
