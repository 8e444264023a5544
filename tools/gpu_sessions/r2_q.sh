#!/bin/bash
# round 2, session Q (2 GPUs): after the hardening of the peer-mailbox set-up / rebalance buffers / NCCL group
# handling: every multi-rank test (host back-end on one GPU, NCCL on two), search tests, short bench at 2 ranks
mkdir -p gpurun_out
PGN_TEST_WORLDS=2 timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_search.py tests/test_gpu_heuristic_kat.py -m gpu -q -x > gpurun_out/r2q_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 5 --warmup 3 --no-side-legs > gpurun_out/r2q_bench_n2.json 2> gpurun_out/r2q_bench_n2.err; echo "bench n=2 rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2q_bench_n2.json') if l.startswith('{')][-1])
print(2, 'value %.4g'%d['value'], 'ms/step %.1f'%d['ms_per_step'], 'parity', d.get('parity_ok'), {k:round(v,1) for k,v in d['stage_ms_per_step'].items()})
PY
