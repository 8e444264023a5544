#!/bin/bash
mkdir -p gpurun_out
B=tools/evalbench/_build
( for f in Oscillatory8 CZero8 Product8 Gaussian8; do for mb in 4 6 7 8; do timeout 120 $B/eb_${f}_mb$mb 5 4; done; done
  for mb in 4 6 7 8; do timeout 120 $B/eb_CZero10_mb$mb 3 4; done
  for mb in 3 4; do timeout 120 $B/eb_Oscillatory8_w8mb$mb 5 4; done ) > gpurun_out/evalbench3.jsonl 2>&1
cat gpurun_out/evalbench3.jsonl | cut -c1-140
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_search.py -m gpu -q -x > gpurun_out/pytest_multi2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_multi2.log
tail -15 gpurun_out/pytest_multi2.log
for n in 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 2 --warmup 1 --no-side-legs > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
echo "bench n=$n rc=$?"; tail -3 gpurun_out/bench_n$n.err; cut -c1-400 gpurun_out/bench_n$n.json
done
timeout 600 python bench.py --steps 2 --warmup 1 --no-side-legs > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cut -c1-400 gpurun_out/bench_n1.json
