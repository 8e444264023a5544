#!/bin/bash
# round 2, session L (2 GPUs): where the threshold-search kernel spends its time (cycle counters in k_search),
# at 1 and 2 ranks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_heuristic_kat.py -m gpu -q -x 2>&1 | tail -3
PGN_SEARCH_DEBUG=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-side-legs > gpurun_out/r2l_bench_n1_debug.json 2> gpurun_out/r2l_bench_n1_debug.err; echo "bench n=1 rc=$?"
PGN_SEARCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 1 --warmup 3 --no-side-legs > gpurun_out/r2l_bench_n2_debug.json 2> gpurun_out/r2l_bench_n2_debug.err; echo "bench n=2 rc=$?"
for n in 1 2; do
grep "pgn device search" gpurun_out/r2l_bench_n${n}_debug.err | tail -42 | python -c "
import sys,re
rows=[]
for l in sys.stdin:
    m=re.search(r'n_local=(\d+) probes=(\d+) computed=(\d+) peer exchanges=(\d+) cycles: kernel=(-?\d+) scans=(-?\d+) exchanges=(-?\d+) elements scanned=(\d+)', l)
    if m: rows.append([int(x) for x in m.groups()])
print('ranks $n: searches', len(rows))
for r in rows[:21]:
    print('  n_local %9d probes %4d computed %3d exchanges %3d  kernel %7.1f us  scans %7.1f us  exchanges %7.1f us  scanned %9d' % (r[0], r[1], r[2], r[3], r[4]/1965., r[5]/1965., r[6]/1965., r[7]))
"
done
