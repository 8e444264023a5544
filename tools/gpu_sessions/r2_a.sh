#!/bin/bash
# round 2, session A: issue-slot microbenchmark + golden trajectories of every BASELINE configuration from the
# unmodified reference binary (oracle/_ref/ref_pagani)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
timeout 120 tools/evalbench/_build/issue_mix > gpurun_out/r2a_issue_mix.jsonl 2>&1; echo "issue_mix rc=$?"
cat gpurun_out/r2a_issue_mix.jsonl
rm -rf gpurun_out/golden
timeout 1500 python tools/make_golden_trajectories.py --out gpurun_out/golden > gpurun_out/r2a_golden.log 2>&1; echo "golden rc=$?"
tail -30 gpurun_out/r2a_golden.log
rm -rf gpurun_out/golden/scratch
du -sh gpurun_out/golden/*
