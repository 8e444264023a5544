#!/bin/bash
# one GPU call: GPU tests, probe timings, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep 'Model name' >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/probe.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?" >> gpurun_out/bench.err
tail -25 gpurun_out/pytest_gpu.log
tail -12 gpurun_out/probe.log
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
