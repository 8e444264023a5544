#!/bin/bash
# first GPU call: parity (quick) + probe timings
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python tools/parity_sweep.py --quick --out gpurun_out/parity > gpurun_out/parity_quick.log 2>&1
echo "parity rc=$?" >> gpurun_out/parity_quick.log
timeout 900 python tools/gpu_probe.py > gpurun_out/probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/probe.log
tail -20 gpurun_out/parity_quick.log
tail -20 gpurun_out/probe.log
