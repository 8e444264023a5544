#!/bin/bash
mkdir -p gpurun_out
B=tools/evalbench/_build
( for exe in eb_Oscillatory8_old eb_Oscillatory8_mb1 eb_Oscillatory8_unitw eb_Oscillatory8_w8 eb_Oscillatory8_w2 eb_Product8_old eb_Product8_mb1; do
    timeout 120 $B/$exe 5 4; done ) > gpurun_out/evalbench2.jsonl 2>&1
cat gpurun_out/evalbench2.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 2 -c 1 \
  -o gpurun_out/eval_f1_8d_new -f $B/eb_Oscillatory8_mb1 4 3 > gpurun_out/ncu_f1_new.log 2>&1
echo "ncu f1 new rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:evaluate_kernel -s 2 -c 1 \
  -o gpurun_out/eval_prod8_new -f $B/eb_Product8_mb1 4 3 > gpurun_out/ncu_prod_new.log 2>&1
echo "ncu prod new rc=$?"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
