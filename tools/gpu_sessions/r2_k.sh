#!/bin/bash
# round 2, session K (8 GPUs): NCCL + in-kernel NVLink exchanges at 4 and 8 ranks against the reference goldens
# (incl. BASELINE config 5, F_4_8D at 1e-9), then the bench at 8, 4 and 1 ranks
mkdir -p gpurun_out
PGN_TEST_WORLDS=4,8 timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -q -x -k "golden" > gpurun_out/r2k_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2k_pytest_multi.log
port=29520
for n in 8 4 1; do
  port=$((port+1))
  if [ $n = 1 ]; then
    timeout 300 python bench.py --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2k_bench_n$n.json 2> gpurun_out/r2k_bench_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2k_bench_n$n.json 2> gpurun_out/r2k_bench_n$n.err
  fi
  echo "bench n=$n rc=$?"
done
python - <<'PY'
import json
base=None
for n in (1, 4, 8):
    try:
        d=json.loads([l for l in open('gpurun_out/r2k_bench_n%d.json' % n) if l.startswith('{')][-1])
        if n == 1: base = d['value']
        print(n, 'value %.4g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'eff %.3f' % (d['value']/(n*base)) if base else '', 'e2e %.4g'%d['e2e']['value'], 'parity', d.get('parity_ok'), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()}, 'launches', d['gpu_launches'])
    except Exception as e: print(n, 'no line', e)
PY
tail -3 gpurun_out/r2k_bench_n8.err | cut -c1-300
