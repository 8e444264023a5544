#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
# launch list of one bench step (cold-cache, serialised: shares only)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
  python bench.py --steps 1 --warmup 0 --no-side-legs > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu launches rc=$?"
# full capture of the evaluation kernel: F_1_8D and F_4_8D on 4^8 = 65536 regions
timeout 900 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 1 -c 1 \
  -o gpurun_out/eval_f1_8d -f python tools/ncu_eval.py 1 8 4 2 > gpurun_out/ncu_f1.log 2>&1
echo "ncu f1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 1 -c 1 \
  -o gpurun_out/eval_f4_8d -f python tools/ncu_eval.py 4 8 4 2 > gpurun_out/ncu_f4.log 2>&1
echo "ncu f4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_threshold -s 40 -c 1 \
  -o gpurun_out/threshold -f python tools/gpu_probe.py 1,8,1e-5 > gpurun_out/ncu_thr.log 2>&1
echo "ncu thr rc=$?"
ls -la gpurun_out
