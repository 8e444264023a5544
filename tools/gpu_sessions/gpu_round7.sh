#!/bin/bash
# full round check: GPU tests, smoke, bench (all legs), ncu launch list of the bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; tail -4 gpurun_out/bench_n1.err; cut -c1-600 gpurun_out/bench_n1.json
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "bench ref rc=$?"; tail -4 gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-side-legs > gpurun_out/bench_ncu.log 2>&1
echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_bench.csv > gpurun_out/launches_summary.txt 2>&1; cat gpurun_out/launches_summary.txt
