#!/bin/bash
# N GPUs: multi-rank tests, then the bench at 1 and N ranks
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_search.py -m gpu -q -x > gpurun_out/pytest_multi3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi3.log
tail -15 gpurun_out/pytest_multi3.log
for n in 1 $N; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --steps 2 --warmup 3 --no-side-legs > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 2 --warmup 3 --no-side-legs > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  echo "bench n=$n rc=$?"; tail -2 gpurun_out/bench_n$n.err | cut -c1-200
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_n$n.json') if l.startswith('{')][-1])
print($n, 'value %.4g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'launches', d['gpu_launches'], 'roof %.3f'%d['roofline']['frac'])
PY
done
