#!/bin/bash
# bench only at N ranks (scaling insurance run)
N=${1:-4}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3 --no-side-legs > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n=$N rc=$?"; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1])
print($N, 'value %.4g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'launches', d['gpu_launches'], 'roof %.3f'%d['roofline']['frac'])
PY
